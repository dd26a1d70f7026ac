// orb.cuh -- device view of the ORB-compatible ("mode B") extractor's working set.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace slamcu {

constexpr int kMaxLevels = 12;
constexpr int kOrbEdge = 31;        // edgeThreshold
constexpr int kOrbPatch = 31;       // patchSize
constexpr int kOrbHalfPatch = 15;

struct OrbLevel {
    int rows, cols, pitch, mwords;
    size_t off;      // byte offset of this level inside a frame's pyramid block (level 0 lives in SeqView::img)
    size_t moff;     // word offset of this level's corner mask inside a frame's mask block
    size_t coff;     // element offset of this level's candidate lists inside a frame's candidate block
    size_t soff;     // byte offset of this level inside a frame's FAST-score block (rows * pitch bytes per level)
    int capc;        // candidate capacity
    int quota;       // nfeaturesPerLevel
    float scale;     // (float)pow(scaleFactor, level)
    // INTER_LINEAR_EXACT tables for producing this level from level-1 (device pointers; unused for level 0):
    // per destination column / row, (first source index << 8) | 8.8 weight of the second tap.  The second tap
    // is index + 1 (clamped); at the clamped ends OpenCV's weight is 0, so the clamp never changes a result.
    const uint32_t* xt; const uint32_t* yt;
};

struct OrbView {
    int nlevels;
    int fast_threshold;
    OrbLevel lv[kMaxLevels];
    size_t pyr_bytes;     // per frame, levels 1..L-1
    size_t mask_words;    // per frame
    size_t cand_total;    // per frame
    uint8_t* pyr;         // [F][pyr_bytes]
    uint8_t* pyrb;        // [F][pyr_bytes]   blurred levels 1..L-1 (level 0 -> SeqView::blur)
    uint32_t* mask;       // [F][mask_words]  FAST-9 + 3x3 NMS + border-filter survivors, 1 bit / pixel
    uint8_t* fscore;      // [F][score_bytes] FAST score, defined only at the survivors' pixels
    size_t score_bytes;   // per frame
    uint32_t* cxy;        // [F][cand_total]  candidates, raster order per level: (y << 16) | x
    int* cscore;          // [F][cand_total]  FAST score
    uint32_t* sxy;        // [F][cand_total]  after retainBest(2*quota)
    float* sresp;         // [F][cand_total]  Harris response
    uint32_t* fxy;        // [F][cand_total]  after retainBest(quota)
    float* fresp;         // [F][cand_total]
    int* n_cand;          // [F][kMaxLevels]
    int* n_sel;           // [F][kMaxLevels]
    int* n_fin;           // [F][kMaxLevels]
    int* octave;          // [F][cap_kp]
    uint32_t* lxy;        // [F][cap_kp]      level coordinates of the final keypoints
    const int8_t* pattern;  // [512][2] rBRIEF sampling points
    const float2* patf;     // the same points as float2 (x, y)
};

// TMA descriptors of the pyramid levels of one sequence: rank-3 uint8 tensors (x, y, frame) with hardware zero fill
// outside the level, box = the FAST tile with its halo.  valid == false -> the kernels stage tiles with LDG instead.
struct OrbTmaps {
    CUtensorMap fast[kMaxLevels];
    CUtensorMap blur[kMaxLevels];
    bool valid = false;
};
constexpr int kFastBoxW = 160, kFastBoxH = 56;  // FSW x FSH of fast9_mask_kernel
constexpr int kBlurBoxW = 160, kBlurBoxH = 66;  // byte tile of blur7_kernel (128 x 60 outputs + 16 px / 3 row halo)

// aux / ev_fork / ev_join: optional second stream (and two events) on which the blur kernels run concurrently
int launch_orb_extract(const SeqView& s, const OrbView& o, int first, int n, cudaStream_t st, cudaStream_t aux = nullptr,
                       cudaEvent_t ev_fork = nullptr, cudaEvent_t ev_join = nullptr, const OrbTmaps* tmaps = nullptr);

}  // namespace slamcu
