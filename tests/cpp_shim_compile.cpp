// compile-and-link check of the header-only C++ drop-in (no GPU needed: it only exercises the error path)
#include "litiv_b200.hpp"
#include <cstdio>
int main() {
    try {
        lvb::BackgroundSubtractorSuBSENSE s;            // throws without a device: "no CPU fallback"
        std::vector<uint8_t> img(64 * 48 * 3, 7), mask;
        s.initialize(lvb::ImageView(img.data(), 48, 64, 3));
        s.apply(lvb::ImageView(img.data(), 48, 64, 3), mask, 1.0);
        std::printf("ran on GPU: %zu mask bytes\n", mask.size());
    } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    return 0;
}
