// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// The drop-in, end to end, in the reference's own types: this program is compiled from the reference's headers
// (StereoEnergy.h, CostVolumeEnergy.h, LayerManager.h, Proposer.h -- where they lie, see oracle/build_ref.py) plus
// include/CudaCostVolumeEnergy.h, the adapter a maintainer would add.  It drives the loop of
// FastGCStereo::localExpansionMovesForLayer_CPU (FastGCStereo.h:22-72, the doGC == false branch, cells visited
// sequentially so that cv::theRNG is deterministic) and, for EVERY proposal, evaluates the unary costs twice through the
// StereoEnergy virtual interface: once with the reference's CPU energy (CostVolumeEnergy / NaiveStereoEnergy) and once
// with the CUDA energy behind the adapter; cv::Mat views, Plane, Reusable and LayerManager rectangles are the
// reference's.  The fusion step then continues with the CPU costs, so both energies always see the same proposals.
//
//   dropin_check [--naive] [--cpu-self-check | --cpu-float-check] [--threads T] [--W n --H n --D n --K n]
// --threads T > 1 runs the cells of a disjoint group in an OpenMP `parallel for`, as the reference does (FastGCStereo.h:30):
// the virtuals are then called concurrently (cv::theRNG is per thread, so the proposals differ from run to run, but every
// call is still compared with the CPU energy on the same proposal).
// --cpu-self-check replaces the CUDA energy by a second CPU instance (validates this harness without a GPU);
// --cpu-float-check by the reference's own single-precision variant (filterName "GFfloat", FastGuidedImageFilter<float>):
// the size of its deviation from the double filter shows how well conditioned the scene is for any FP32 implementation.
// --batched runs the loop as INTEGRATION.md section 3 restructures it: per (layer, group) one CudaCostVolumeEnergy::GroupPlan,
// per proposal step ONE evaluation of all cells of the group (image-shaped output on even steps, per-cell tiles wrapped in
// cv::Mat headers on odd steps), then the unchanged per-cell fusion; the reference energy still checks every cell.
// Prints one JSON line; exit code 0 iff every call agreed (1e-4 relative, COST_FOR_INVALID masks identical).
#include <opencv2/opencv.hpp>
#include "Utilities.hpp"
#include "Plane.h"
#include "StereoEnergy.h"
#include "CostVolumeEnergy.h"
#include "LayerManager.h"
#include "CudaCostVolumeEnergy.h"
#include <cstring>
#include <omp.h>

namespace {

struct Tally {
    long calls = 0, px = 0, bad = 0, mask_mismatch = 0;
    double worst = 0;  // max |a-b| / (1e-4 * max(|a|, 1e-3))
    double t_ref = 0, t_test = 0;  // seconds this thread spent inside the reference CPU energy / the energy under test
};

void compare(const cv::Mat& a, const cv::Mat& b, const cv::Rect& r, Tally& t) {
    t.calls++;
    for (int y = r.y; y < r.y + r.height; y++) {
        const float* pa = a.ptr<float>(y);
        const float* pb = b.ptr<float>(y);
        for (int x = r.x; x < r.x + r.width; x++) {
            t.px++;
            const bool ia = pa[x] == (float)StereoEnergy::COST_FOR_INVALID, ib = pb[x] == (float)StereoEnergy::COST_FOR_INVALID;
            if (ia != ib) { t.mask_mismatch++; continue; }
            if (ia) continue;
            if (std::isnan(pa[x]) && std::isnan(pb[x])) continue;
            double e = std::fabs((double)pa[x] - pb[x]) / (1e-4 * std::max(std::fabs((double)pa[x]), 1e-3));
            if (!(e <= 1.0)) t.bad++;
            if (e > t.worst || std::isnan(e)) t.worst = e;
        }
    }
}

cv::Mat synthetic_image(int H, int W, uint64_t seed) {
    cv::RNG rng(seed);
    cv::Mat im(H, W, CV_8UC3);
    // smooth blobs + texture, so that the guided filter sees edges
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double s = 0.5 + 0.5 * std::sin(0.11 * x + 0.07 * y) * std::cos(0.05 * x - 0.13 * y);
            int edge = ((x / 23 + y / 17) & 1) ? 60 : 0;
            for (int c = 0; c < 3; c++) {
                int v = (int)(40 + 120 * s + edge + (c * 17) + rng.uniform(0, 30));
                im.at<cv::Vec3b>(y, x)[c] = (uchar)std::min(255, std::max(0, v));
            }
        }
    return im;
}

}  // namespace

int main(int argc, char** argv) {
    int W = 160, H = 120, D = 16, K = 3, windR = 20, threads = 1;
    bool naive = false, self = false, selff = false, batched = false;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--naive")) naive = true;
        else if (!std::strcmp(argv[i], "--cpu-self-check")) self = true;
        else if (!std::strcmp(argv[i], "--cpu-float-check")) self = selff = true;
        else if (!std::strcmp(argv[i], "--batched")) batched = true;
        else if (i + 1 < argc && !std::strcmp(argv[i], "--W")) W = std::atoi(argv[++i]);
        else if (i + 1 < argc && !std::strcmp(argv[i], "--H")) H = std::atoi(argv[++i]);
        else if (i + 1 < argc && !std::strcmp(argv[i], "--D")) D = std::atoi(argv[++i]);
        else if (i + 1 < argc && !std::strcmp(argv[i], "--K")) K = std::atoi(argv[++i]);
        else if (i + 1 < argc && !std::strcmp(argv[i], "--threads")) threads = std::max(1, std::atoi(argv[++i]));
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    try {
        const float maxdisp = (float)D - 1;
        cv::Mat imL = synthetic_image(H, W, 11), imR = synthetic_image(H, W, 12);
        std::vector<float> vL((size_t)D * H * W), vR((size_t)D * H * W);
        cv::RNG vr(99);
        for (auto& v : vL) v = (float)vr;
        for (auto& v : vR) v = (float)vr;
        int sizes[3] = {D, H, W};
        cv::Mat volL(3, sizes, CV_32F, vL.data()), volR(3, sizes, CV_32F, vR.data());

        Parameters param(1.0f, windR, "GF", 0.0001f);  // paramsGF, main.cpp:73
        if (!naive) param.th_col = 0.5f;                // main.cpp:26,351

        Parameters paramB = param;
        if (selff) paramB.filterName = "GFfloat";
        std::unique_ptr<StereoEnergy> A, B;  // A: the reference's CPU energy; B: the energy under test
        if (naive) {
            A = std::make_unique<NaiveStereoEnergy>(imL, imR, param, maxdisp);
            if (self) B = std::make_unique<NaiveStereoEnergy>(imL, imR, paramB, maxdisp);
            else B = std::make_unique<CudaNaiveStereoEnergy>(imL, imR, param, maxdisp);
        } else {
            A = std::make_unique<CostVolumeEnergy>(imL, imR, volL, volR, param, maxdisp);
            if (self) B = std::make_unique<CostVolumeEnergy>(imL, imR, volL, volR, paramB, maxdisp);
            else B = std::make_unique<CudaCostVolumeEnergy>(imL, imR, volL, volR, param, maxdisp);  // main.cpp:386, with the adapter
        }

        LayerManager layermng(W, H, windR, 0);
        const int units[3] = {5, 15, 25};  // main.cpp:304-306
        for (int u : units) layermng.addLayer(u);
        const cv::Rect imageDomain(0, 0, W, H);
        Tally init, moves;

        for (int mode = 0; mode < 2; mode++) {
            cv::theRNG().state = 1234 + mode;
            cv::Mat currentCost(H, W, CV_32F, cv::Scalar(INFINITY));
            cv::Mat currentLabeling(H, W, CV_32FC4, cv::Scalar::all(0));
            cv::Mat costB(H, W, CV_32F, cv::Scalar(INFINITY));
            // initCurrentFast (FastGCStereo.h:95-116): a random label per unit region of layer 0
            auto& layer0 = layermng.layers[0];
            for (size_t j = 0; j < layer0.unitRegions.size(); j++) {
                cv::Rect unit = layer0.unitRegions[j];
                int n = cv::theRNG().uniform(0, unit.height * unit.width);
                cv::Point pnt(unit.x + n % unit.width, unit.y + n / unit.width);
                Plane label = A->createRandomLabel(pnt);
                currentLabeling(unit) = label.toScalar();
                cv::Rect filterRegion = cv::Rect(unit.x - windR, unit.y - windR, unit.width + windR * 2, unit.height + windR * 2) & imageDomain;
                StereoEnergy::Reusable ra, rb;
                A->ComputeUnaryPotential(filterRegion, unit, currentCost(filterRegion), label, ra, mode);
                B->ComputeUnaryPotential(filterRegion, unit, costB(filterRegion), label, rb, mode);
                compare(currentCost, costB, unit, init);
            }
            // local expansion moves (FastGCStereo.h:22-72), doGC == false
            for (int iteration = 0; iteration < 2; iteration++)
                for (auto& layer : layermng.layers) {
                    cv::Mat proposalCost(H, W, CV_32F), proposalCostB(H, W, CV_32F);
                    for (size_t j = 0; j < layer.disjointRegionSets.size() && batched && !self; j++) {
                        // ---- one evaluation per proposal step for all cells of the group -----------------
                        const std::vector<int>& cells = layer.disjointRegionSets[j];
                        const int nc = (int)cells.size();
                        std::vector<cv::Rect> fr(nc), tr(nc);
                        for (int n = 0; n < nc; n++) { fr[n] = layer.filterRegions[cells[n]]; tr[n] = layer.sharedRegions[cells[n]]; }
                        CudaCostVolumeEnergy::GroupPlan plan(*static_cast<CudaCostVolumeEnergy*>(B.get()), fr, tr);  // once per (layer, group) in a real run
                        std::vector<float> tiles(plan.tileFloats());
                        std::vector<StereoEnergy::Reusable> reus(nc);
                        ExpansionProposer p1(1);
                        RandomProposer p2(K, maxdisp);
                        IProposer* protos[2] = {&p1, &p2};
                        int step = 0;
                        for (IProposer* proto : protos) {
                            std::vector<IProposer*> prop(nc);
                            for (int n = 0; n < nc; n++) { prop[n] = proto->createInstance(); prop[n]->startIterations(currentLabeling, layer.unitRegions[cells[n]], iteration); }
                            while (prop[0]->isContinued()) {  // all cells of a group run the same number of steps (same K, same outerIter)
                                std::vector<Plane> labels(nc);
                                for (int n = 0; n < nc; n++) labels[n] = prop[n]->getNextProposal();
                                const bool as_tiles = (step++ & 1) != 0;
                                const double t1 = omp_get_wtime();
                                if (as_tiles) plan.evaluateTiles(labels, tiles.data(), mode, true);
                                else plan.evaluate(labels, proposalCostB, mode, true);
                                moves.t_test += omp_get_wtime() - t1;
                                for (int n = 0; n < nc; n++) {  // reference energy per cell + the unchanged fusion
                                    const cv::Rect& sharedRegion = tr[n];
                                    const double t0 = omp_get_wtime();
                                    A->ComputeUnaryPotential(fr[n], sharedRegion, proposalCost(fr[n]), labels[n], reus[n], mode);
                                    moves.t_ref += omp_get_wtime() - t0;
                                    if (as_tiles) plan.tile(tiles.data(), n).copyTo(proposalCostB(sharedRegion));
                                    compare(proposalCost, proposalCostB, sharedRegion, moves);
                                    cv::Mat subCurrentCost = currentCost(sharedRegion), subProposalCost = proposalCost(sharedRegion);
                                    cv::Mat subCurrentLabeling = currentLabeling(sharedRegion);
                                    cv::Mat updateMask = subCurrentCost > subProposalCost;
                                    subProposalCost.copyTo(subCurrentCost, updateMask);
                                    subCurrentLabeling.setTo(labels[n].toScalar(), updateMask);
                                }
                            }
                            for (IProposer* p : prop) delete p;
                        }
                    }
                    for (size_t j = 0; j < layer.disjointRegionSets.size() && !(batched && !self); j++) {
                        std::vector<Tally> part(threads);
#pragma omp parallel for num_threads(threads) if (threads > 1)
                        for (int n = 0; n < (int)layer.disjointRegionSets[j].size(); n++) {
                            Tally& moves = part[omp_get_thread_num()];
                            int r = layer.disjointRegionSets[j][n];
                            auto& sharedRegion = layer.sharedRegions[r];
                            auto& unitRegion = layer.unitRegions[r];
                            cv::Mat subCurrentCost = currentCost(sharedRegion);
                            cv::Mat subProposalCost = proposalCost(sharedRegion);
                            cv::Mat subCurrentLabeling = currentLabeling(sharedRegion);
                            StereoEnergy::Reusable reusable, reusableB;
                            ExpansionProposer p1(1);
                            RandomProposer p2(K, maxdisp);
                            IProposer* protos[2] = {&p1, &p2};
                            for (IProposer* proto : protos) {
                                IProposer* prop = proto->createInstance();
                                prop->startIterations(currentLabeling, unitRegion, iteration);
                                while (prop->isContinued()) {
                                    Plane label = prop->getNextProposal();
                                    const double t0 = omp_get_wtime();
                                    A->ComputeUnaryPotential(layer.filterRegions[r], sharedRegion, proposalCost(layer.filterRegions[r]), label, reusable, mode);
                                    const double t1 = omp_get_wtime();
                                    B->ComputeUnaryPotential(layer.filterRegions[r], sharedRegion, proposalCostB(layer.filterRegions[r]), label, reusableB, mode);
                                    moves.t_ref += t1 - t0; moves.t_test += omp_get_wtime() - t1;
                                    compare(proposalCost, proposalCostB, sharedRegion, moves);
                                    cv::Mat updateMask = subCurrentCost > subProposalCost;
                                    subProposalCost.copyTo(subCurrentCost, updateMask);
                                    subCurrentLabeling.setTo(label.toScalar(), updateMask);
                                }
                                delete prop;
                            }
                        }
                        for (const Tally& t : part) {
                            moves.calls += t.calls; moves.px += t.px; moves.bad += t.bad; moves.mask_mismatch += t.mask_mismatch;
                            moves.t_ref += t.t_ref; moves.t_test += t.t_test;
                            if (t.worst > moves.worst || std::isnan(t.worst)) moves.worst = t.worst;
                        }
                    }
                }
        }
        const long px = init.px + moves.px, bad = init.bad + moves.bad, mm = init.mask_mismatch + moves.mask_mismatch;
        const double worst = std::max(init.worst, moves.worst);
        // both energies: no pixel out of tolerance (the CUDA NaiveStereoEnergy repeats getAffineTransform's LU + warpAffine's inversion)
        const bool ok = mm == 0 && bad == 0;
        // thread-seconds inside the two energies during the expansion moves (same calls, same threads): how much faster the
        // unchanged loop gets its unary costs through the adapter
        char timing[160];
        std::snprintf(timing, sizeof timing, "\"thread_seconds_reference_energy\": %.3f, \"thread_seconds_energy_under_test\": %.3f, ", moves.t_ref, moves.t_test);
        char extra[128] = "";
        if (!self) {  // how many of the concurrent calls the library served with combined launches
            int64_t batches = 0, calls = 0;
            lexp_combine_stats(static_cast<CudaCostVolumeEnergy*>(B.get())->context(), &batches, &calls);
            std::snprintf(extra, sizeof extra, "\"combined_launches\": %lld, \"combined_calls\": %lld, ", (long long)batches, (long long)calls);
        }
        std::printf("{\"energy\": \"%s\", \"under_test\": \"%s\", \"W\": %d, \"H\": %d, \"D\": %d, \"init_calls\": %ld, \"move_calls\": %ld, "
                    "\"pixels\": %ld, \"out_of_tolerance\": %ld, \"mask_mismatch\": %ld, \"worst_err_over_tol\": %.4g, \"threads\": %d, %s%s\"ok\": %s}\n",
                    naive ? "NaiveStereoEnergy" : "CostVolumeEnergy", selff ? "cpu GFfloat" : self ? "cpu-self-check" : batched ? "CudaCostVolumeEnergy::GroupPlan (batched loop)" : "CudaCostVolumeEnergy adapter", W, H, D,
                    init.calls, moves.calls, px, bad, mm, worst, threads, extra, timing, ok ? "true" : "false");
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 3;
    }
}
