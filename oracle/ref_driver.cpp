// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// C entry points around the reference's OWN classes, compiled from the reference's headers where they lie
// (/root/reference/LocalExpansionStereo, read-only; see oracle/build_ref.py) over the mini cv:: layer in
// oracle/cvshim/.  Nothing here restates the algorithm: every cost, mask, statistic, rectangle and random label
// returned below is produced by CostVolumeEnergy / NaiveStereoEnergy / FastGuidedImageFilter<double> / LayerManager /
// RandomProposer / StereoEnergy::createRandomLabel themselves.  The output, oracle/_ref/liblexp_ref.so, pins the
// restated oracles (oracle/lexp_oracle.py, oracle/lexp_oracle.c) and generates tests/golden/ref_*.npz.
#include <opencv2/opencv.hpp>
#include "Utilities.hpp"
#include "Plane.h"
#include "StereoEnergy.h"
#include "CostVolumeEnergy.h"
#include "LayerManager.h"
#include "Proposer.h"
#include "FastGCStereo.h"
#include <omp.h>
#include <malloc.h>

namespace {

// the reference's optimiser class with its protected graph-cut move and state opened up (nothing overridden)
struct GCProbe : FastGCStereo {
    using FastGCStereo::FastGCStereo;
    using FastGCStereo::expansionMoveBK;
    using PMStereoBase::currentCost_;
    using PMStereoBase::currentLabeling_;
    using PMStereoBase::currentLabeling_m_;
};

struct ref_ctx {
    std::unique_ptr<StereoEnergy> own;   // owner of the energy until the optimiser probe takes it over (setStereoEnergyCPU)
    StereoEnergy* energy = nullptr;
    std::unique_ptr<GCProbe> gc;
    cv::Mat im[2];
    int H, W, D, kind;
    float max_disp, min_disp;
};
struct SmoothProbe : StereoEnergy {
    static const std::vector<cv::Mat>& coeff_of(const StereoEnergy& e, int m) { return (e.*(&SmoothProbe::smoothnessCoeff))[m]; }
};

// protected members of the reference's classes, reached through pointers to members named via a derived class
struct EnergyProbe : CostVolumeEnergy {
    static const std::unique_ptr<IJointFilter>& filter_of(const CostVolumeEnergy& e, int m) { return (e.*(&EnergyProbe::filter))[m]; }
};
struct NaiveProbe : NaiveStereoEnergy {
    static const std::unique_ptr<IJointFilter>& filter_of(const NaiveStereoEnergy& e, int m) { return (e.*(&NaiveProbe::filter))[m]; }
    static const cv::Mat& exi_of(const NaiveStereoEnergy& e, int m) { return (e.*(&NaiveProbe::ExI))[m]; }
};
struct FilterProbe : FastGuidedImageFilter<double> {
    typedef GuidedImageFilter<double> G;
    static void planes(const G& f, const cv::Mat* out[9]) {
        out[0] = &(f.*(&FilterProbe::mean_I_r)); out[1] = &(f.*(&FilterProbe::mean_I_g)); out[2] = &(f.*(&FilterProbe::mean_I_b));
        out[3] = &(f.*(&FilterProbe::invrr)); out[4] = &(f.*(&FilterProbe::invrg)); out[5] = &(f.*(&FilterProbe::invrb));
        out[6] = &(f.*(&FilterProbe::invgg)); out[7] = &(f.*(&FilterProbe::invgb)); out[8] = &(f.*(&FilterProbe::invbb));
    }
};

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
int ref_max_threads() { return omp_get_max_threads(); }

// kind 0: CostVolumeEnergy (volumes [D][H][W] float, borrowed -- the caller keeps them alive); kind 1: NaiveStereoEnergy.
// Images are 8-bit BGR [H][W][3].  filterName "GF" (FastGuidedImageFilter<double>), as main.cpp selects for both modes.
void* ref_create(int kind, int H, int W, int D, const uchar* imL, const uchar* imR, float* volL, float* volR, int windR, float eps,
                 float th_col, float th_grad, float alpha, float max_disp, float min_disp) {
    try {
        // every cv::Mat temporary of the filter is a malloc: keep large blocks in the per-thread arenas instead of mmap/munmap,
        // which would serialise the OpenMP threads in the kernel
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        auto c = std::make_unique<ref_ctx>();
        c->H = H; c->W = W; c->D = D; c->kind = kind; c->max_disp = max_disp; c->min_disp = min_disp;
        c->im[0] = cv::Mat(H, W, CV_8UC3, (void*)imL).clone();
        c->im[1] = cv::Mat(H, W, CV_8UC3, (void*)imR).clone();
        Parameters params(20.f, windR, "GF", eps);
        params.th_col = th_col; params.th_grad = th_grad; params.alpha = alpha;
        if (kind == 0) {
            int sz[3] = {D, H, W};
            cv::Mat vL(3, sz, CV_32F, volL), vR(3, sz, CV_32F, volR);
            c->own = std::make_unique<CostVolumeEnergy>(c->im[0], c->im[1], vL, vR, params, max_disp, min_disp);
        } else {
            c->own = std::make_unique<NaiveStereoEnergy>(c->im[0], c->im[1], params, max_disp, min_disp);
        }
        c->energy = c->own.get();
        return c.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_destroy(void* h) { delete (ref_ctx*)h; }

// One call of the virtual, exactly as FastGCStereo.h:47 issues it: `out` is the filterRect-sized view
// proposalCost(filterRect) ([fh][fw] float, caller-initialised; only the target sub-rectangle is written).
int ref_unary(void* h, const int frect[4], const int trect[4], const float plane[4], int mode, int with_check, float* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        cv::Rect fr(frect[0], frect[1], frect[2], frect[3]), tr(trect[0], trect[1], trect[2], trect[3]);
        cv::Mat costs(fr.height, fr.width, CV_32F, out);
        Plane p(plane[0], plane[1], plane[2], plane[3]);
        StereoEnergy::Reusable reusable;
        if (with_check) c->energy->ComputeUnaryPotential(fr, tr, costs, p, reusable, mode);
        else c->energy->ComputeUnaryPotentialWithoutCheck(fr, tr, costs, p, reusable, mode);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// The reference's inner loop over one disjoint group (FastGCStereo.h:30-49) without the fusion step: cells in an
// OpenMP parallel for, one Reusable per cell kept across its K proposals, each proposal evaluated into the shared
// H x W `cost_image` through the view cost_image(filterRect).  planes: [n][K][4].
int ref_unary_group(void* h, int n, const int* frects, const int* trects, const float* planes, int K, int mode, int with_check, float* cost_image, int nthreads) {
    ref_ctx* c = (ref_ctx*)h;
    cv::Mat proposalCost(c->H, c->W, CV_32F, cost_image);
    int failed = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        try {
            cv::Rect fr(frects[4 * i], frects[4 * i + 1], frects[4 * i + 2], frects[4 * i + 3]), tr(trects[4 * i], trects[4 * i + 1], trects[4 * i + 2], trects[4 * i + 3]);
            StereoEnergy::Reusable reusable;
            for (int k = 0; k < K; k++) {
                const float* pl = planes + ((size_t)i * K + k) * 4;
                Plane label(pl[0], pl[1], pl[2], pl[3]);
                if (with_check) c->energy->ComputeUnaryPotential(fr, tr, proposalCost(fr), label, reusable, mode);
                else c->energy->ComputeUnaryPotentialWithoutCheck(fr, tr, proposalCost(fr), label, reusable, mode);
            }
        } catch (const std::exception&) {
#pragma omp atomic
            failed++;
        }
    }
    if (failed) g_err = "ref_unary_group: a cell threw";
    return failed;
}

// guided-filter statistics of view `mode`: 9 planes [H][W] double: mean_I r,g,b then inv rr,rg,rb,gg,gb,bb
int ref_stats(void* h, int mode, double* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        const IJointFilter* f = c->kind == 0 ? EnergyProbe::filter_of(*static_cast<CostVolumeEnergy*>(c->energy), mode).get()
                                             : NaiveProbe::filter_of(*static_cast<NaiveStereoEnergy*>(c->energy), mode).get();
        auto g = dynamic_cast<const GuidedImageFilter<double>*>(f);
        if (!g) throw std::runtime_error("not a guided filter");
        const cv::Mat* pl[9];
        FilterProbe::planes(*g, pl);
        for (int k = 0; k < 9; k++)
            for (int y = 0; y < c->H; y++) std::memcpy(out + ((size_t)k * c->H + y) * c->W, pl[k]->ptr<double>(y), sizeof(double) * c->W);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// NaiveStereoEnergy::ExI[mode] as [H][W][4] float
int ref_exi(void* h, int mode, float* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        if (c->kind != 1) throw std::runtime_error("ExI exists for NaiveStereoEnergy only");
        const cv::Mat& e = NaiveProbe::exi_of(*static_cast<NaiveStereoEnergy*>(c->energy), mode);
        for (int y = 0; y < c->H; y++) std::memcpy(out + (size_t)y * c->W * 4, e.ptr<float>(y), sizeof(float) * 4 * c->W);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// StereoEnergy::IsValiLabel(plane, rect) -> [h][w] uchar (255 / 0)
int ref_valid_mask(void* h, const float plane[4], const int rect[4], uchar* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        cv::Mat m = c->energy->IsValiLabel(Plane(plane[0], plane[1], plane[2], plane[3]), cv::Rect(rect[0], rect[1], rect[2], rect[3]));
        for (int y = 0; y < m.rows; y++) std::memcpy(out + (size_t)y * m.cols, m.ptr<uchar>(y), m.cols);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// ---- cv::theRNG-driven label generation ----------------------------------------------------------------------------
void ref_rng_seed(uint64_t s) { cv::theRNG().state = s; }
uint64_t ref_rng_state() { return cv::theRNG().state; }
void ref_create_random_label(void* h, int x, int y, float out[4]) {
    Plane p = ((ref_ctx*)h)->energy->createRandomLabel(cv::Point(x, y));
    out[0] = p.a; out[1] = p.b; out[2] = p.c; out[3] = p.v;
}
// RandomProposer over `unit` of the labeling [H][W][4] (Plane per pixel), FastGCStereo.h:39-46: returns the proposals
// generated while isContinued(); out [K][4]
int ref_random_proposals(float* labeling, int H, int W, const int unit[4], int outerIter, int K, float max_disp, float min_disp, float* out) {
    cv::Mat lab(H, W, CV_32FC4, labeling);
    RandomProposer proto(K, max_disp, min_disp);
    IProposer* prop = proto.createInstance();
    prop->startIterations(lab, cv::Rect(unit[0], unit[1], unit[2], unit[3]), outerIter);
    int n = 0;
    while (prop->isContinued()) {
        Plane p = prop->getNextProposal();
        out[4 * n] = p.a; out[4 * n + 1] = p.b; out[4 * n + 2] = p.c; out[4 * n + 3] = p.v;
        n++;
    }
    delete prop;
    return n;
}
// ---- the PatchMatch phase of one disjoint group: the body of FastGCStereo::localExpansionMovesForLayer_CPU with
// doGC == false (FastGCStereo.h:30-61), re-typed here because FastGCStereo.h needs the un-vendored maxflow sources.  The
// proposers are the reference's own ExpansionProposer / RandomProposer instances, created and driven exactly as at :41-46
// (createInstance, startIterations, isContinued, getNextProposal); the energy is the reference's.  Two harness additions:
// (1) cv::theRNG().state is set to states[n][k] right before proposal k of cell n is drawn (the reference leaves the
// assignment of random streams to cells to the OpenMP runtime); (2) a replay proposer for the slots whose planes the
// caller supplies (kind 0: fixed plane lists, the RansacProposer slot).
namespace {
class ListProposer : public IProposer {
    const float* planes_;   // [K][4] of the current cell
public:
    ListProposer(int K, const float* planes) : IProposer(K), planes_(planes) {}
    IProposer* createInstance() override { return new ListProposer(K, planes_); }
    void startIterations(cv::Mat, cv::Rect, int) override { iter = 0; }
    Plane getNextProposal() override { const float* p = planes_ + 4 * iter++; return Plane(p[0], p[1], p[2], p[3]); }
    bool isContinued() override { return iter < K; }
};
}  // namespace
// cells: n; rects as int[4] each; prop_kind[j] / prop_K[j]: the layer's proposer list (0 list, 1 ExpansionProposer, 2 RandomProposer);
// list_planes: [n][total list steps][4]; states: [n][max_steps] cv::RNG states; cur_cost [H][W], cur_label [H][W][4] updated in
// place; planes_out [n][max_steps][4] receives every proposal; steps_out[n] their count.  Returns < 0 on error.
int ref_pm_group(void* h, int mode, int n, const int* units, const int* shareds, const int* filts, int nprop, const int* prop_kind,
                 const int* prop_K, int outer_iter, float max_disp, float min_disp, const float* list_planes, int list_steps,
                 const uint64_t* states, int max_steps, float* cur_cost, float* cur_label, float* planes_out, int* steps_out, int nthreads) {
    ref_ctx* c = (ref_ctx*)h;
    cv::Mat currentCost(c->H, c->W, CV_32F, cur_cost), currentLabeling(c->H, c->W, CV_32FC4, cur_label);
    cv::Mat proposalCost = cv::Mat(c->H, c->W, CV_32F);                                                  // :25
    int failed = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        try {
            cv::Rect unitRegion(units[4 * i], units[4 * i + 1], units[4 * i + 2], units[4 * i + 3]);
            cv::Rect sharedRegion(shareds[4 * i], shareds[4 * i + 1], shareds[4 * i + 2], shareds[4 * i + 3]);
            cv::Rect filterRegion(filts[4 * i], filts[4 * i + 1], filts[4 * i + 2], filts[4 * i + 3]);
            cv::Mat subCurrentCost = currentCost(sharedRegion);                                           // :36-38
            cv::Mat subProposalCost = proposalCost(sharedRegion);
            cv::Mat subCurrentLabeling = currentLabeling(sharedRegion);
            StereoEnergy::Reusable reusable;
            int step = 0, list_at = 0;
            for (int j = 0; j < nprop; j++) {                                                             // :41
                IProposer* prop;
                if (prop_kind[j] == 1) { ExpansionProposer proto(prop_K[j]); prop = proto.createInstance(); }
                else if (prop_kind[j] == 2) { RandomProposer proto(prop_K[j], max_disp, min_disp); prop = proto.createInstance(); }
                else { ListProposer proto(prop_K[j], list_planes + ((size_t)i * list_steps + list_at) * 4); prop = proto.createInstance(); list_at += prop_K[j]; }
                prop->startIterations(currentLabeling, unitRegion, outer_iter);                           // :44
                while (prop->isContinued()) {                                                             // :45
                    if (step >= max_steps) throw std::runtime_error("more proposals than max_steps");
                    cv::theRNG().state = states[(size_t)i * max_steps + step];
                    Plane label = prop->getNextProposal();                                                // :47
                    float* po = planes_out + ((size_t)i * max_steps + step) * 4;
                    po[0] = label.a; po[1] = label.b; po[2] = label.c; po[3] = label.v;
                    c->energy->ComputeUnaryPotential(filterRegion, sharedRegion, proposalCost(filterRegion), label, reusable, mode);   // :49
                    cv::Mat updateMask = subCurrentCost > subProposalCost;                                // :56
                    subProposalCost.copyTo(subCurrentCost, updateMask);                                   // :58
                    subCurrentLabeling.setTo(label.toScalar(), updateMask);                               // :59
                    step++;
                }
                delete prop;
            }
            steps_out[i] = step;
        } catch (const std::exception& e) {
#pragma omp critical
            { g_err = e.what(); failed++; }
        }
    }
    return failed ? -failed : 0;
}
// ---- pairwise terms and the graph-cut move ----------------------------------------------------------
// The smoothness parameters live in the energy's public `params`; the coefficient maps are rebuilt by the reference's own
// initSmoothnessCoeff() (StereoEnergy.h:131-163).
int ref_set_smoothness(void* h, float lambda, float omega, float th_smooth, float epsilon) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        c->energy->params.lambda = lambda; c->energy->params.omega = omega;
        c->energy->params.th_smooth = th_smooth; c->energy->params.epsilon = epsilon;
        c->energy->initSmoothnessCoeff();
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// smoothnessCoeff[mode][k] without its 1-pixel margin: out [8][H][W]
int ref_smooth_coeff(void* h, int mode, float* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        const std::vector<cv::Mat>& co = SmoothProbe::coeff_of(*c->energy, mode);
        for (int k = 0; k < (int)co.size() && k < 8; k++)
            for (int y = 0; y < c->H; y++)
                memcpy(out + ((size_t)k * c->H + y) * c->W, co[k].ptr<float>(y + 1) + 1, (size_t)c->W * sizeof(float));
        return (int)co.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// computeSmoothnessTermsExpansion(labeling0_m, label1, region, cost00, cost01, cost10, onlyForward = true, mode) (StereoEnergy.h:398-453),
// called as expansionMoveBK does (FastGCStereo.h:422).  labeling [H][W][4]; the 1-pixel margin is zero (PMStereoBase.h:44-45).
// out [3][8][rh][rw]: cost00 / cost01 / cost10 for the neighbours the reference fills (forward ones: NB_GE 1, NB_EG 3, NB_LG 6, NB_GG 7).
int ref_smooth_terms_expansion(void* h, int mode, const float* labeling, const float plane[4], const int region[4], float* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        cv::Mat lab_m = cv::Mat::zeros(c->H + 2, c->W + 2, CV_32FC4);
        cv::Mat(c->H, c->W, CV_32FC4, (void*)labeling).copyTo(lab_m(cv::Rect(1, 1, c->W, c->H)));
        std::vector<cv::Mat> c00, c01, c10;
        const cv::Rect rg(region[0], region[1], region[2], region[3]);
        c->energy->computeSmoothnessTermsExpansion(lab_m, Plane(plane[0], plane[1], plane[2], plane[3]), rg, c00, c01, c10, true, mode);
        const size_t n = (size_t)rg.width * rg.height;
        memset(out, 0, 3 * 8 * n * sizeof(float));
        const std::vector<cv::Mat>* all[3] = {&c00, &c01, &c10};
        for (int t = 0; t < 3; t++)
            for (int k = 0; k < (int)all[t]->size() && k < 8; k++) {
                const cv::Mat& m = (*all[t])[k];
                if (m.empty()) continue;
                for (int y = 0; y < rg.height; y++) memcpy(out + ((size_t)t * 8 + k) * n + (size_t)y * rg.width, m.ptr<float>(y), (size_t)rg.width * sizeof(float));
            }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// One disjoint group of FastGCStereo::localExpansionMovesForLayer_CPU with doGC == true (FastGCStereo.h:30-61): the loop body is
// re-typed as in ref_pm_group (same harness additions: per-(cell, step) cv::RNG states, list proposer), but the move itself is the
// reference's OWN FastGCStereo::expansionMoveBK (FastGCStereo.h:411-597) -- graph construction, pairwise terms, boundary terms --
// running on the reference's own state members, over oracle/maxflow/graph.h (the un-vendored BK library's interface, see there).
// flows_out [n][max_steps]: the value expansionMoveBK returns (= energy of the move's minimum cut); masks are applied as at :58-59.
int ref_gc_group(void* h, int mode, int n, const int* units, const int* shareds, const int* filts, int nprop, const int* prop_kind,
                 const int* prop_K, int outer_iter, const float* list_planes, int list_steps, const uint64_t* states, int max_steps,
                 float* cur_cost, float* cur_label, float* planes_out, int* steps_out, double* flows_out, int nthreads) {
    ref_ctx* c = (ref_ctx*)h;
    try {
        if (!c->gc) {   // the reference's optimiser object; it takes the energy over exactly as main.cpp:386 does
            Parameters prm = c->energy->params;
            c->gc = std::make_unique<GCProbe>(c->im[0], c->im[1], prm, c->max_disp, c->min_disp);
            if (c->own) c->gc->setStereoEnergyCPU(std::move(c->own));
        }
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    GCProbe& gc = *c->gc;
    cv::Mat(c->H, c->W, CV_32F, cur_cost).copyTo(gc.currentCost_[mode]);
    cv::Mat(c->H, c->W, CV_32FC4, cur_label).copyTo(gc.currentLabeling_[mode]);   // the view into currentLabeling_m_ (margin stays zero)
    cv::Mat currentCost = gc.currentCost_[mode], currentLabeling = gc.currentLabeling_[mode];
    cv::Mat proposalCost = cv::Mat(c->H, c->W, CV_32F);                                                  // :25
    int failed = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        try {
            cv::Rect unitRegion(units[4 * i], units[4 * i + 1], units[4 * i + 2], units[4 * i + 3]);
            cv::Rect sharedRegion(shareds[4 * i], shareds[4 * i + 1], shareds[4 * i + 2], shareds[4 * i + 3]);
            cv::Rect filterRegion(filts[4 * i], filts[4 * i + 1], filts[4 * i + 2], filts[4 * i + 3]);
            cv::Mat subCurrentCost = currentCost(sharedRegion);                                           // :36-38
            cv::Mat subProposalCost = proposalCost(sharedRegion);
            cv::Mat subCurrentLabeling = currentLabeling(sharedRegion);
            StereoEnergy::Reusable reusable;
            int step = 0, list_at = 0;
            for (int j = 0; j < nprop; j++) {                                                             // :41
                IProposer* prop;
                if (prop_kind[j] == 1) { ExpansionProposer proto(prop_K[j]); prop = proto.createInstance(); }
                else if (prop_kind[j] == 2) { RandomProposer proto(prop_K[j], c->max_disp, c->min_disp); prop = proto.createInstance(); }
                else { ListProposer proto(prop_K[j], list_planes + ((size_t)i * list_steps + list_at) * 4); prop = proto.createInstance(); list_at += prop_K[j]; }
                prop->startIterations(currentLabeling, unitRegion, outer_iter);                           // :44
                while (prop->isContinued()) {                                                             // :45
                    if (step >= max_steps) throw std::runtime_error("more proposals than max_steps");
                    cv::theRNG().state = states[(size_t)i * max_steps + step];
                    Plane label = prop->getNextProposal();                                                // :47
                    float* po = planes_out + ((size_t)i * max_steps + step) * 4;
                    po[0] = label.a; po[1] = label.b; po[2] = label.c; po[3] = label.v;
                    c->energy->ComputeUnaryPotential(filterRegion, sharedRegion, proposalCost(filterRegion), label, reusable, mode);   // :49
                    cv::Mat updateMask = cv::Mat_<uchar>(sharedRegion.size());                            // :53
                    const double flow = gc.expansionMoveBK(updateMask, label, sharedRegion, subProposalCost, mode);   // :54
                    if (flows_out) flows_out[(size_t)i * max_steps + step] = flow;
                    subProposalCost.copyTo(subCurrentCost, updateMask);                                   // :58
                    subCurrentLabeling.setTo(label.toScalar(), updateMask);                               // :59
                    step++;
                }
                delete prop;
            }
            steps_out[i] = step;
        } catch (const std::exception& e) {
#pragma omp critical
            { g_err = e.what(); failed++; }
        }
    }
    gc.currentCost_[mode].copyTo(cv::Mat(c->H, c->W, CV_32F, cur_cost));
    gc.currentLabeling_[mode].copyTo(cv::Mat(c->H, c->W, CV_32FC4, cur_label));
    return failed ? -failed : 0;
}
// the BK stand-in (oracle/maxflow/graph.h) on a grid graph given as arrays: tr [h][w] net terminal capacities, cap [4][h][w] forward
// arcs (GE, EG, LG, GG); mask[s] = what_segment(s) == SOURCE; returns the flow of the residual network (no add_tweights constants:
// positive tr is added as source weight, negative as sink weight)
double shim_grid_mincut(int w, int h, const float* tr, const float* cap, uchar* mask) {
    typedef Graph<float, float, double> G;
    G g(w * h, 4 * w * h);
    g.add_node(w * h);
    const int dx[4] = {1, 0, -1, 1}, dy[4] = {0, 1, 1, 1};
    for (int s = 0; s < w * h; s++) g.add_tweights(s, tr[s] > 0 ? tr[s] : 0.f, tr[s] < 0 ? -tr[s] : 0.f);
    for (int d = 0; d < 4; d++)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int xx = x + dx[d], yy = y + dy[d];
                if (xx >= 0 && xx < w && yy < h) g.add_edge(y * w + x, yy * w + xx, cap[((size_t)d * h + y) * w + x], 0);
            }
    const double f = g.maxflow();
    for (int s = 0; s < w * h; s++) mask[s] = g.what_segment(s) == G::SOURCE;
    return f;
}
// ---- file formats either side of the path: the reference's own cvutils::io functions --------------------
int ref_save_pfm(const char* path, const float* img, int H, int W) {
    try { cvutils::io::save_pfm_file(path, cv::Mat(H, W, CV_32F, (void*)img)); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int ref_read_pfm(const char* path, float* out, int H, int W) {
    try {
        cv::Mat m = cvutils::io::read_pfm_file(path);
        if (m.empty() || m.rows != H || m.cols != W || m.channels() != 1) { g_err = "read_pfm_file: absent or wrong shape"; return -1; }
        for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, m.ptr<float>(y), (size_t)W * sizeof(float));
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// loadMatBinary(path, vol, readHeader = false) into a D x H x W float volume, as main.cpp:353-358
int ref_load_acrt(const char* path, float* out, int D, int H, int W) {
    try {
        int sizes[] = {D, H, W};
        cv::Mat vol(3, sizes, CV_32F, out);
        return cvutils::io::loadMatBinary(path, vol, false) ? 0 : -1;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// StereoEnergy::computeDisparities(labeling) (StereoEnergy.h:269-272); labeling [H][W][4], out [H][W]
int ref_disparities(void* h, const float* labeling, float* out) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        cv::Mat d = c->energy->computeDisparities(cv::Mat(c->H, c->W, CV_32FC4, (void*)labeling));
        for (int y = 0; y < c->H; y++) memcpy(out + (size_t)y * c->W, d.ptr<float>(y), (size_t)c->W * sizeof(float));
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// total energy of a labeling as the reference evaluates it: sum of currentCost + computeSmoothnessCost (StereoEnergy.h:165-199)
double ref_smoothness_cost(void* h, int mode, const float* labeling) {
    try {
        ref_ctx* c = (ref_ctx*)h;
        cv::Mat lab_m = cv::Mat::zeros(c->H + 2, c->W + 2, CV_32FC4);
        cv::Mat(c->H, c->W, CV_32FC4, (void*)labeling).copyTo(lab_m(cv::Rect(1, 1, c->W, c->H)));
        return c->energy->computeSmoothnessCost(lab_m, mode);
    } catch (const std::exception& e) { g_err = e.what(); return -1.0; }
}
// initCurrentFast with a given label per unit region (FastGCStereo.h:101-113; the random draw of :105-106 is the caller's):
// currentLabeling(unit) = label; ComputeUnaryPotential(unit +- windR, unit, currentCost(filterRegion), label)
int ref_pm_init(void* h, int mode, int n, const int* units, const float* labels, int windR, float* cur_cost, float* cur_label, int nthreads) {
    ref_ctx* c = (ref_ctx*)h;
    cv::Mat currentCost(c->H, c->W, CV_32F, cur_cost), currentLabeling(c->H, c->W, CV_32FC4, cur_label);
    const cv::Rect imageDomain(0, 0, c->W, c->H);
    int failed = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for
    for (int j = 0; j < n; j++) {
        try {
            cv::Rect unit(units[4 * j], units[4 * j + 1], units[4 * j + 2], units[4 * j + 3]);
            Plane label(labels[4 * j], labels[4 * j + 1], labels[4 * j + 2], labels[4 * j + 3]);
            currentLabeling(unit) = label.toScalar();                                                     // :107
            const int R = windR;
            cv::Rect filterRegion = cv::Rect(unit.x - R, unit.y - R, unit.width + R * 2, unit.height + R * 2) & imageDomain;   // :110
            StereoEnergy::Reusable reusable;
            c->energy->ComputeUnaryPotential(filterRegion, unit, currentCost(filterRegion), label, reusable, mode);            // :111
        } catch (const std::exception& e) {
#pragma omp critical
            { g_err = e.what(); failed++; }
        }
    }
    return failed ? -failed : 0;
}

void ref_plane_normal(const float plane[4], float out[3]) {
    cv::Vec<float, 3> n = Plane(plane[0], plane[1], plane[2], plane[3]).GetNormal();
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
}
void ref_create_plane(const float n[3], float z, float x, float y, float v, float out[4]) {
    Plane p = Plane::CreatePlane(n[0], n[1], n[2], z, x, y, v);
    out[0] = p.a; out[1] = p.b; out[2] = p.c; out[3] = p.v;
}

// ---- LayerManager::addLayer ---------------------------------------------------------------------------------------------
void* ref_layer_create(int W, int H, int windR, int unit) {
    auto lm = new LayerManager(W, H, windR, 0);
    lm->addLayer(unit);
    return lm;
}
void ref_layer_destroy(void* h) { delete (LayerManager*)h; }
void ref_layer_counts(void* h, int out[4]) {
    auto& L = ((LayerManager*)h)->layers[0];
    out[0] = L.heightBlocks; out[1] = L.widthBlocks; out[2] = (int)L.unitRegions.size(); out[3] = (int)L.disjointRegionSets.size();
}
void ref_layer_rects(void* h, int* unit, int* shared, int* filter) {
    auto& L = ((LayerManager*)h)->layers[0];
    for (size_t i = 0; i < L.unitRegions.size(); i++) {
        const cv::Rect* r[3] = {&L.unitRegions[i], &L.sharedRegions[i], &L.filterRegions[i]};
        int* o[3] = {unit, shared, filter};
        for (int k = 0; k < 3; k++) { o[k][4 * i] = r[k]->x; o[k][4 * i + 1] = r[k]->y; o[k][4 * i + 2] = r[k]->width; o[k][4 * i + 3] = r[k]->height; }
    }
}
int ref_layer_group(void* h, int g, int* idx) {
    auto& S = ((LayerManager*)h)->layers[0].disjointRegionSets[g];
    if (idx) for (size_t i = 0; i < S.size(); i++) idx[i] = S[i];
    return (int)S.size();
}

// ---- the cv:: layer's own primitives, exported so that tests can hold them against the real OpenCV (cv2) ---------------
int shim_box_sum(const void* src, int is_double, int h, int w, int R, void* dst) {
    try {
        int t = is_double ? CV_64FC1 : CV_32FC1;
        cv::Mat s(h, w, t, (void*)src), d(h, w, t, dst);
        cv::boxFilter(s, d, -1, cv::Size(2 * R + 1, 2 * R + 1), cv::Point(-1, -1), false, cv::BORDER_CONSTANT);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int shim_get_affine(const float src[6], const float dst[6], double M[6]) {
    cv::Point2f s[3], d[3];
    for (int i = 0; i < 3; i++) { s[i] = cv::Point2f(src[2 * i], src[2 * i + 1]); d[i] = cv::Point2f(dst[2 * i], dst[2 * i + 1]); }
    cv::Mat m = cv::getAffineTransform(s, d);
    for (int i = 0; i < 6; i++) M[i] = m.at<double>(i / 3, i % 3);
    return 0;
}
int shim_warp_affine(float* src, int H, int W, int cn, double M[6], int h, int w, float* dst) {
    try {
        cv::Mat s(H, W, CV_MAKETYPE(CV_32F, cn), src), d(h, w, CV_MAKETYPE(CV_32F, cn), dst), m(2, 3, CV_64FC1, M);
        cv::warpAffine(s, d, m, cv::Size(w, h), cv::INTER_LINEAR, cv::BORDER_REPLICATE);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int shim_bgr2gray(float* src, int H, int W, float* dst) {
    try { cv::Mat s(H, W, CV_32FC3, src), d(H, W, CV_32FC1, dst); cv::cvtColor(s, d, cv::COLOR_BGR2GRAY); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int shim_sobel_x(float* src, int H, int W, double scale, float* dst) {
    try { cv::Mat s(H, W, CV_32FC1, src), d(H, W, CV_32FC1, dst); cv::Sobel(s, d, CV_32F, 1, 0, 1, scale, 0, cv::BORDER_REPLICATE); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"
