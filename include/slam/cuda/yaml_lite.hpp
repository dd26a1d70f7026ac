// yaml_lite.hpp -- reader for the OpenCV-FileStorage YAML files the reference's constructors parse: flat `key: value`
// scalars (feature_detector.hpp:54-94, feature_matcher.cpp:19-59), flow sequences (`ImageSize: [w, h]`) and
// `!!opencv-matrix` nodes (camera.yml: K<i>, D<i>; common.hpp:76-122).  Used by the CUDA adapters so that they do not
// need OpenCV just to read a dozen numbers; where OpenCV is available cv::FileStorage gives the same values.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace slam::cuda {

class YamlLite {
public:
    explicit YamlLite(const std::string& path) {
        std::ifstream in(path);
        m_open = in.good();
        std::string line, last;
        while (std::getline(in, line)) {
            std::string s;
            char quote = 0;
            for (char c : line) {  // strip comments outside quotes
                if (quote) { if (c == quote) quote = 0; }
                else if (c == '"' || c == '\'') quote = c;
                else if (c == '#') break;
                s.push_back(c);
            }
            if (!s.empty() && (s[0] == ' ' || s[0] == '\t')) {  // continuation of the current node (opencv-matrix body)
                if (!last.empty()) m_values[last] += " " + trim(s);
                continue;
            }
            if (s.empty() || s[0] == '%' || s.rfind("---", 0) == 0) continue;
            const size_t colon = s.find(':');
            if (colon == std::string::npos) continue;
            std::string key = trim(s.substr(0, colon)), val = trim(s.substr(colon + 1));
            if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front())
                val = val.substr(1, val.size() - 2);
            m_values[key] = val;
            last = key;
        }
    }
    bool isOpened() const { return m_open; }
    bool has(const std::string& key) const { return m_values.count(key) != 0; }
    // cv::FileNode >> int / float / string semantics: a missing node leaves 0 / 0.f / ""; a real read as int rounds
    int getInt(const std::string& key) const {
        auto it = m_values.find(key);
        if (it == m_values.end()) return 0;
        char* end = nullptr;
        const double v = std::strtod(it->second.c_str(), &end);
        return end == it->second.c_str() ? 0 : static_cast<int>(v < 0 ? v - 0.5 : v + 0.5);
    }
    float getFloat(const std::string& key) const {
        auto it = m_values.find(key);
        return it == m_values.end() ? 0.0F : static_cast<float>(std::strtod(it->second.c_str(), nullptr));
    }
    // the numbers of a flow sequence `[a, b, ...]`, or of the `data: [...]` array of an !!opencv-matrix node
    std::vector<double> getDoubles(const std::string& key) const {
        std::vector<double> out;
        auto it = m_values.find(key);
        if (it == m_values.end()) return out;
        const std::string& v = it->second;
        size_t from = v.find("data:");
        from = v.find('[', from == std::string::npos ? 0 : from);
        const size_t to = v.find(']', from == std::string::npos ? 0 : from);
        if (from == std::string::npos || to == std::string::npos) return out;
        const char* p = v.c_str() + from + 1;
        const char* end = v.c_str() + to;
        while (p < end) {
            char* q = nullptr;
            const double d = std::strtod(p, &q);
            if (q == p) { p++; continue; }
            out.push_back(d);
            p = q;
        }
        return out;
    }
    std::string getString(const std::string& key) const {
        auto it = m_values.find(key);
        return it == m_values.end() ? std::string() : it->second;
    }

private:
    static std::string trim(const std::string& s) {
        const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    bool m_open = false;
    std::map<std::string, std::string> m_values;
};

}  // namespace slam::cuda
