// example_stereo.cpp -- the reference's call sequence (src/estimator/estimator.rs:85-89,190) through the
// C++ mirror: construct once, process_frame per stereo pair, read the features from the Frame.
// Build: g++ -std=c++17 example_stereo.cpp -L../lib -lrsvio_b200 -Wl,-rpath,'$ORIGIN/../lib' -o example_stereo
#include <cstdio>
#include <vector>

#include "rsvio/feature_tracker.hpp"

struct Frame {  // stand-in for src/estimator/frame.rs
    std::vector<rsvio::Feature> left_features, right_features;
    void add_left_feature(const rsvio::Feature &f) { left_features.push_back(f); }
    void add_right_feature(const rsvio::Feature &f) { right_features.push_back(f); }
};

int main() {
    const std::uint32_t w = 752, h = 480;
    std::vector<std::uint8_t> left(w * h), right(w * h);
    for (std::uint32_t y = 0; y < h; ++y)
        for (std::uint32_t x = 0; x < w; ++x) {  // blocky texture, right image shifted by 8 px
            left[y * w + x] = (std::uint8_t)((((x / 12) * 37 + (y / 12) * 91) % 200) + 20);
            right[y * w + x] = (std::uint8_t)(((((x + 8) / 12) * 37 + (y / 12) * 91) % 200) + 20);
        }
    try {
        rsvio::StereoPatchTracker<6> tracker(50, 20, 0.01);  // estimator.rs:85-89 with config/euroc_vio.yaml:42-45
        for (int t = 0; t < 3; ++t) {
            Frame frame;
            tracker.process_frame({left.data(), w, h}, {right.data(), w, h}, frame);
            std::printf("frame %d: %zu left / %zu right features\n", t, frame.left_features.size(), frame.right_features.size());
        }
        auto pts = tracker.get_track_points();
        std::printf("tracked: %zu / %zu\n", pts.first.size(), pts.second.size());
    } catch (const rsvio::Error &e) {
        std::fprintf(stderr, "rsvio error %d: %s\n", (int)e.status, e.what());
        return e.status == RSVIO_ERR_CUDA ? 2 : 1;  // no GPU: fails loudly, never falls back
    }
    return 0;
}
