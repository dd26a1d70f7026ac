set -e
cd $GRAFT_REPO_ROOT 2>/dev/null || true
sed 's/MaxFeatures: 2000/MaxFeatures: 1000/' test/data/feature_detector_orb.yml > test/data/_det_tum.yml
sed 's/MaxFeatures: 2000/MaxFeatures: 10000/' test/data/feature_detector_orb.yml > test/data/_det_4k.yml
B=slam_cin0051_b200/build/slam_bench
echo "config2 (C++ driver)"; $B -c test/data/feature_detector_orb.yml -m test/data/feature_matcher_orb.yml -f 1000 -s 3 -K 2560 -C 500
echo "config3 TUM-shape 640x480, 1000 kp, match + RANSAC"; $B -c test/data/_det_tum.yml -m test/data/feature_matcher_orb.yml -W 640 -H 480 -f 1000 -s 3 -K 1280 -C 500 -p 17 -e 525,525,319.5,239.5
echo "config3 without RANSAC"; $B -c test/data/_det_tum.yml -m test/data/feature_matcher_orb.yml -W 640 -H 480 -f 1000 -s 3 -K 1280 -C 500 -p 17
echo "config4 4K 3840x2160, 10000 kp"; $B -c test/data/_det_4k.yml -m test/data/feature_matcher_orb.yml -W 3840 -H 2160 -f 48 -s 3 -K 12288 -C 24 -p 28
rm -f test/data/_det_tum.yml test/data/_det_4k.yml
