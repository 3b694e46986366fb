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
    try { lvb::BackgroundSubtractorPAWCS p; p.refreshModel(1, 0.0f); } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    try {
        lvb::BackgroundSubtractorViBe_3ch v; lvb::BackgroundSubtractorPBAS_1ch b;   // throw without a device
        std::vector<uint8_t> img(64 * 48 * 3, 7), mask;
        v.initialize(lvb::ImageView(img.data(), 48, 64, 3)); v.apply(lvb::ImageView(img.data(), 48, 64, 3), mask);
        b.initialize(lvb::ImageView(img.data(), 48, 64, 1)); b.apply(lvb::ImageView(img.data(), 48, 64, 1), mask);
        std::printf("ViBe / PBAS ran on GPU\n");
    } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    try {
        lvb::EdgeDetectorLBSP e;                        // throws without a device
        std::vector<uint8_t> img(64 * 48 * 3, 7), edges(64 * 48);
        e.apply_threshold(lvb::ImageView(img.data(), 48, 64, 3), edges.data(), e.getDefaultThreshold());
        std::printf("EdgeDetectorLBSP ran on GPU\n");
    } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); }
    {   // features2d/test/lbsp.cpp:4-19 on the shim (host only)
        lvb::LBSP l(size_t(20));
        bool threw = false;
        try { l.borderSize(2); } catch(const lvb::Exception&) { threw = true; }
        bool threw2 = false;
        try { lvb::LBSP bad(-0.5f, size_t(0)); } catch(const lvb::Exception&) { threw2 = true; }
        std::printf("LBSP invariants %s\n", (threw && threw2 && l.windowSize() / 2 == l.borderSize() && l.borderSize(1) == 2 && l.descriptorSize() == 2 && l.descriptorType() == 2 && l.defaultNorm() == 6) ? "ok" : "BROKEN");
    }
    lvb::BinClassif bc; bc.nTP = 6; bc.nTN = 80; bc.nFP = 4; bc.nFN = 10;
    const lvb::BinClassifMetrics m(bc);   // host arithmetic only
    std::printf("F-measure %.6f total %llu\n", m.dFMeasure, (unsigned long long)bc.total());
    return 0;
}
