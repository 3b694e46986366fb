// drives the header-only C++ drop-in (include/litiv_b200.hpp) on the GPU: the reference's sample loop
// (samples/changedet/src/main.cpp:33-74: initialize with the first frame, apply with learning rate 1 for the first frames, then
// the default) on frames read from a raw file; writes the final masks to a raw file so that the test can compare them with the
// Python path. usage: shim_gpu <algo 0|1|2|3 ViBe|4 PBAS> <w> <h> <c> <nframes> <in.raw> <out.raw>
#include "litiv_b200.hpp"
#include <cstdio>
#include <cstdlib>
#include <memory>
int main(int argc, char** argv) {
    if(argc != 8) return 2;
    const int algo = atoi(argv[1]), w = atoi(argv[2]), h = atoi(argv[3]), c = atoi(argv[4]), n = atoi(argv[5]);
    std::vector<uint8_t> frames((size_t)w * h * c * n), masks((size_t)w * h * (n - 1));
    FILE* f = fopen(argv[6], "rb");
    if(!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
    fclose(f);
    const size_t fs0 = (size_t)w * h * c;
    if(algo == 3 || algo == 4) { // ViBe / PBAS: plain cv::BackgroundSubtractor shape (initialize(img), apply with the class default)
        try {
            std::vector<uint8_t> bg(fs0);
            if(algo == 3) {
                lvb::BackgroundSubtractorViBe_3ch v(20, 20, 2, 0, /*seed*/ 5);
                v.initialize(lvb::ImageView(frames.data(), h, w, c));
                for(int t = 1; t < n; ++t) v.apply(lvb::ImageView(frames.data() + fs0 * t, h, w, c), masks.data() + (size_t)w * h * (t - 1));
                v.getBackgroundImage(bg.data());
            } else {
                lvb::BackgroundSubtractorPBAS_3ch v(30, 16.0f, 35, 2, 0, /*seed*/ 5);
                v.initialize(lvb::ImageView(frames.data(), h, w, c));
                for(int t = 1; t < n; ++t) v.apply(lvb::ImageView(frames.data() + fs0 * t, h, w, c), masks.data() + (size_t)w * h * (t - 1));
                v.getBackgroundImage(bg.data());
            }
            std::printf("ran on GPU: %d frames, bg[0]=%d\n", n - 1, (int)bg[0]);
        } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); return 1; }
        f = fopen(argv[7], "wb");
        if(!f || fwrite(masks.data(), 1, masks.size(), f) != masks.size()) return 4;
        fclose(f);
        return 0;
    }
    try {
        std::unique_ptr<lvb::SubtractorBase> p;
        if(algo == 0) p.reset(new lvb::BackgroundSubtractorLOBSTER(4, 30, 35, 2, 0, 0.333f, 0, /*seed*/ 5));
        else if(algo == 1) p.reset(new lvb::BackgroundSubtractorSuBSENSE(3, 30, 50, 2, 100, 0.333f, 0, /*seed*/ 5));
        else p.reset(new lvb::BackgroundSubtractorPAWCS(2, 20, 50, 100, 0.333f, 0, /*seed*/ 5));
        const size_t fs = (size_t)w * h * c;
        p->initialize(lvb::ImageView(frames.data(), h, w, c));
        for(int t = 1; t < n; ++t) {
            const double lr = algo == 0 ? p->getDefaultLearningRate() : (t <= 5 ? 1.0 : p->getDefaultLearningRate());
            p->apply(lvb::ImageView(frames.data() + fs * t, h, w, c), masks.data() + (size_t)w * h * (t - 1), lr);
        }
        std::vector<uint8_t> bg(fs);
        p->getBackgroundImage(bg.data());
        lvb::BinClassif bc;
        bc.accumulate(*p, lvb::ImageView(masks.data() + (size_t)w * h * (n - 2), h, w, 1)); // the last mask scored against itself
        const lvb::BinClassifMetrics m(bc);
        std::printf("ran on GPU: %d frames, F-measure vs itself %.3f, bg[0]=%d\n", n - 1, m.dFMeasure, (int)bg[0]);
    } catch(const lvb::Exception& e) { std::printf("lvb::Exception: %s\n", e.what()); return 1; }
    f = fopen(argv[7], "wb");
    if(!f || fwrite(masks.data(), 1, masks.size(), f) != masks.size()) return 4;
    fclose(f);
    return 0;
}
