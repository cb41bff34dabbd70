// Compile-and-link check of include/CudaCostVolumeEnergy.h against the cv stub and liblexp_cuda.so.
#define LEXP_ADAPTER_NO_REFERENCE_INCLUDES
#include "cv_stub.h"
#include "../../include/CudaCostVolumeEnergy.h"
#include <cstdio>
int main(int argc, char**) {
    std::printf("lexp version %d\n", lexp_version());
    if (argc > 100) {  // never executed: only instantiates the adapter so that every call is type-checked and linked
        cv::Mat im, vol;
        CudaCostVolumeEnergy e(im, im, vol, vol, Parameters(), 63.f);
        cv::Rect r{0, 0, 1, 1};
        StereoEnergy::Reusable ru;
        e.ComputeUnaryPotential(r, r, im, Plane{0, 0, 0, 0}, ru, 0);
        e.ComputeUnaryPotentialWithoutCheck(r, r, im, Plane{0, 0, 0, 0}, ru, 0);
        std::vector<cv::Rect> rs; std::vector<Plane> ps;
        e.ComputeUnaryPotentialBatch(rs, rs, im, ps);
        CudaCostVolumeEnergy::GroupPlan gp(e, rs, rs);
        gp.evaluate(ps, im, 0, true);
        gp.evaluateTiles(ps, nullptr, 1, false);
        (void)gp.cells(); (void)gp.tileFloats();
        CudaCostVolumeEnergy::PatchMatchPhase pm(e, 0, {5, 15}, {{{LEXP_PROP_EXPANSION, 1}, {LEXP_PROP_RANDOM, 8}}, {{LEXP_PROP_EXPANSION, 3}}});
        pm.begin(); pm.begin(im, im); pm.init(ps); (void)pm.iteration(0, 1234u); pm.get(im, im);
        if (e.failed()) e.throwIfFailed();
        CudaNaiveStereoEnergy n(im, im, Parameters(), 63.f);
        n.ComputeUnaryPotential(r, r, im, Plane{0, 0, 0, 0}, ru, 1);
    }
    return 0;
}
