// lexp_kernels.cuh -- hand-written sm_100a kernels of the unary-cost hot path.
//
// K0  lexp_stats_*      one-time guided-filter statistics       (GuidedFilter.h:58-102)
// K1+K2+K3 fused        lexp_fused_kernel: plane-cost sampling + truncation
//                       (CostVolumeEnergy.h:69-98), guided filter filter_raw
//                       (GuidedFilter.h:142-247) with the sub-region window counts of
//                       createSubregionFilter (GuidedFilter.h:301-326), validity mask
//                       (StereoEnergy.h:577-610, CostVolumeEnergy.h:176-183).
// File:line citations are relative to /root/reference/LocalExpansionStereo/.
//
// Design: one CTA = one output tile of one (cell, plane) call.  The CTA
// streams top-to-bottom over the rows of the tile's dependency cone (tile +- 2R):
//   V-phase (thread = column): gather p, running column sums of {p, I*p} (stage 1) and of
//            {a, b} (stage 2) kept in registers, the 2R+1 rows they will subtract later
//            kept in a thread-private shared-memory ring; epilogue q = (Bb + Ba.I)/N.
//   H-phase (thread = run of 8 columns): horizontal window sums over the rows of the chunk.
// Two CTA barriers per chunk of CH rows; intermediates never leave the SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lexp {

constexpr int kThreads = 128;          // CTA size == max virtual tile width (ow + 4R)
constexpr int kRun = 8;                // columns per H-phase task
constexpr float kCostInvalid = 1000000.0f;  // StereoEnergy.h:45

struct __align__(16) Item {  // one CTA work item (64 B)
    int fx, fy, fw, fh;      // filterRect of the call
    int ox0, oy0, ow, oh;    // output tile (image coordinates), inside targetRect
    int call;                // index of the call (plane / compact slot)
    int compact_off;         // float offset of the tile's first pixel in the compact output
    int compact_stride;      // = targetRect.width
    int flags;               // bit0: targetRect is 1x1 (IsValiLabel fast path, StereoEnergy.h:579-583)
    int pad[4];
};

struct Plane4 { float a, b, c, v; };

struct KParams {
    const float* __restrict__ vol;      // float[D][H][W]
    const uchar4* __restrict__ guide;   // uchar4[H][W] = (c0,c1,c2,0), OpenCV BGR order
    const float* __restrict__ stats;    // float[9][H][W]: mean c0,c1,c2, inv 00,01,02,11,12,22
    const Item* __restrict__ items;
    const Plane4* __restrict__ planes;
    float* __restrict__ out;
    long long out_pitch;                // floats per row (image mode)
    int out_compact;                    // 1: per-call contiguous tiles, 0: H x W image
    int H, W, D;
    float th_col, min_disp, max_disp;
    int with_check;
    int R;                              // guided-filter box radius (windR / 2)
};

__host__ __device__ __forceinline__ int sidx(int x) { return x + (x >> 3); }  // bank-conflict padding
__host__ __device__ __forceinline__ int srow_stride(int vw) {
    int s = sidx(vw + kRun + 7) + 1;
    return s + ((10 - (s & 7)) & 7);  // == 2 (mod 8) in float4 units
}
// dynamic shared memory (bytes) of the fused kernel for a tile of virtual width vw
__host__ __device__ __forceinline__ size_t fused_smem_bytes(int vw, int R, int CH) {
    const int K = 2 * R + 1;
    return (size_t)(K * vw + K * (vw - 2 * R) + 4 * CH * srow_stride(vw)) * sizeof(float4);
}

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// Plane disparity at (X, y): two separately rounded operations each, like the reference's
// `d_base = b*y + c; d = a*x + d_base` compiled without FMA contraction (CostVolumeEnergy.h:73,76).
__device__ __forceinline__ float plane_d(float a, float d_base, int X) {
    return __fadd_rn(__fmul_rn(a, (float)X), d_base);
}

// Classify d (CostVolumeEnergy.h:78-92).  Returns the interpolation weight f1 >= 0 and the two
// slice indices, or f1 = -1 (clamped: C = V[i0]) or f1 = -2 (C = COST_FOR_INVALID, nothing to load).
__device__ __forceinline__ float classify_d(float d, float minD, float maxD, int D0, int D, int& i0, int& i1) {
    if (d < minD) { i0 = i1 = 0; return -1.f; }
    if (d >= maxD) { i0 = i1 = D - 1; return -1.f; }
    if (isnan(d) || isinf(d)) { i0 = i1 = 0; return -2.f; }
    const int d0 = (int)d + D0;  // int(d): truncation toward zero (:83)
    if (d0 + 1 >= D || d0 < 0) { i0 = i1 = 0; return -2.f; }  // (:87-90)
    i0 = d0; i1 = d0 + 1;
    return d - floorf(d);        // (:85)
}

template <int R_T, int CH>
__global__ void __launch_bounds__(kThreads) lexp_fused_kernel(const KParams P) {
    const int R = R_T > 0 ? R_T : P.R;
    const int K = 2 * R + 1;
    extern __shared__ float4 smem[];

    const Item it = P.items[blockIdx.x];
    const Plane4 pl = P.planes[it.call];
    const int t = threadIdx.x;
    const int VW = it.ow + 4 * R, VH = it.oh + 4 * R;
    const int X0 = it.ox0 - 2 * R, Y0 = it.oy0 - 2 * R;
    const int W2 = VW - 2 * R;            // stage-2 (a,b) columns
    const int SW = srow_stride(VW);
    float4* ring1 = smem;                  // [K][VW]   raw {p, I0 p, I1 p, I2 p} rows
    float4* ring2 = ring1 + K * VW;        // [K][W2]   {a0, a1, a2, b} rows
    float4* hbuf1 = ring2 + K * W2;        // [CH][SW]  stage-1 column sums of the chunk
    float4* hout1 = hbuf1 + CH * SW;       // [CH][SW]  stage-1 box sums
    float4* hbuf2 = hout1 + CH * SW;       // [CH][SW]  stage-2 column sums
    float4* hout2 = hbuf2 + CH * SW;       // [CH][SW]  stage-2 box sums

    {   // zero-fill (box filter is zero padded, GuidedFilter.h:43 BORDER_CONSTANT)
        const int total = K * VW + K * W2 + 4 * CH * SW;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = t; i < total; i += kThreads) smem[i] = z;
    }

    const int fx1 = it.fx + it.fw, fy1 = it.fy + it.fh;
    const size_t HW = (size_t)P.H * P.W;
    const int D0 = (int)(-P.min_disp);
    const float th = P.th_col;

    // column roles of this thread
    const int XA = X0 + t;                       // stage-1 gather column
    const bool colA = (t < VW) && XA >= it.fx && XA < fx1;
    const int XC = X0 + R + t;                   // (a,b) column
    const bool colC = (t < W2) && XC >= it.fx && XC < fx1;
    const int XE = it.ox0 + t;                   // output column
    const bool colE = t < it.ow;
    float inv_nxC = 0.f, inv_nxE = 0.f;          // 1 / (#columns of the window inside filterRect), GuidedFilter.h:324
    if (colC) inv_nxC = 1.0f / (float)(min(XC + R, fx1 - 1) - max(XC - R, it.fx) + 1);
    if (colE) inv_nxE = 1.0f / (float)(min(XE + R, fx1 - 1) - max(XE - R, it.fx) + 1);

    const int nChunks = (VH + CH - 1) / CH;
    const int n1 = (W2 + kRun - 1) / kRun;       // stage-1 runs per row
    const int n2 = (it.ow + kRun - 1) / kRun;    // stage-2 runs per row
    const int T1 = ((CH * n1 + 31) / 32) * 32;   // stage-1 tasks padded to a warp boundary
    const int Ttot = T1 + CH * n2;

    float4 acc1 = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc1;
    int slot1 = 0, slot2 = 0;

    // prefetch registers for the gather of one chunk
    float pv0[CH], pv1[CH], pf1[CH];
    uint32_t pg[CH];

    auto prefetch = [&](int chunk) {
#pragma unroll
        for (int r = 0; r < CH; r++) {
            const int v = chunk * CH + r;
            const int y = Y0 + v;
            pf1[r] = -3.f;  // outside filterRect: contributes zero
            pv0[r] = 0.f; pv1[r] = 0.f; pg[r] = 0u;
            if (colA && v < VH && y >= it.fy && y < fy1) {
                const float d_base = __fadd_rn(__fmul_rn(pl.b, (float)y), pl.c);
                const float d = plane_d(pl.a, d_base, XA);
                int i0, i1;
                const float f1 = classify_d(d, P.min_disp, P.max_disp, D0, P.D, i0, i1);
                pf1[r] = f1;
                const size_t pix = (size_t)y * P.W + XA;
                if (f1 > -2.f) pv0[r] = __ldg(P.vol + (size_t)i0 * HW + pix);
                if (f1 >= 0.f) pv1[r] = __ldg(P.vol + (size_t)i1 * HW + pix);
                pg[r] = __ldg(reinterpret_cast<const unsigned int*>(P.guide) + pix);
            }
        }
    };

    prefetch(0);
    __syncthreads();

    for (int itn = 0; itn < nChunks + 2; itn++) {
        // ------------------------------------------------------------------ V-phase
        // (A) stage-1 rows of chunk itn: consume the prefetched samples
        if (itn < nChunks) {
            float cv0[CH], cv1[CH], cf1[CH];
            uint32_t cg[CH];
#pragma unroll
            for (int r = 0; r < CH; r++) { cv0[r] = pv0[r]; cv1[r] = pv1[r]; cf1[r] = pf1[r]; cg[r] = pg[r]; }
            if (itn + 1 < nChunks) prefetch(itn + 1);
            if (t < VW) {
#pragma unroll
                for (int r = 0; r < CH; r++) {
                    float4 nw = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float f1 = cf1[r];
                    if (f1 > -3.f) {
                        float C;
                        if (f1 >= 0.f) C = __fadd_rn(__fmul_rn(1.0f - f1, cv0[r]), __fmul_rn(f1, cv1[r]));  // (:92)
                        else if (f1 > -2.f) C = cv0[r];                                                      // (:78-79)
                        else C = kCostInvalid;                                                               // (:80,:89)
                        const float p = (th < C) ? th : C;                                                   // std::min (:96)
                        const float s = 1.0f / 255.0f;
                        const float i0 = (float)(cg[r] & 0xffu) * s, i1 = (float)((cg[r] >> 8) & 0xffu) * s,
                                    i2 = (float)((cg[r] >> 16) & 0xffu) * s;
                        nw = make_float4(p, i0 * p, i1 * p, i2 * p);                                         // GuidedFilter.h:151-169
                    }
                    float4* slot = ring1 + slot1 * VW + t;
                    const float4 old = *slot;
                    *slot = nw;
                    acc1 = f4sub(f4add(acc1, nw), old);
                    hbuf1[r * SW + sidx(t)] = acc1;  // column sum centred on row v - R
                    slot1 = (slot1 + 1 == K) ? 0 : slot1 + 1;
                }
            }
        }
        // (C) (a,b) rows from the stage-1 box sums of chunk itn-1
        if (itn >= 1 && itn - 1 < nChunks && t < W2) {
#pragma unroll
            for (int r = 0; r < CH; r++) {
                const int v = (itn - 1) * CH + r;
                if (v >= 2 * R && v < VH) {
                    const int yc = Y0 + v - R;
                    float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (colC && yc >= it.fy && yc < fy1) {
                        const float inv_ny = 1.0f / (float)(min(yc + R, fy1 - 1) - max(yc - R, it.fy) + 1);
                        const float invN = inv_nxC * inv_ny;
                        const float4 B = hout1[r * SW + sidx(t + R)];
                        const float* st = P.stats + (size_t)yc * P.W + XC;
                        const float m0 = __ldg(st), m1 = __ldg(st + HW), m2 = __ldg(st + 2 * HW);
                        const float i00 = __ldg(st + 3 * HW), i01 = __ldg(st + 4 * HW), i02 = __ldg(st + 5 * HW);
                        const float i11 = __ldg(st + 6 * HW), i12 = __ldg(st + 7 * HW), i22 = __ldg(st + 8 * HW);
                        const float mp = B.x * invN;                       // GuidedFilter.h:206
                        const float c0 = fmaf(B.y, invN, -m0 * mp);        // :212-214
                        const float c1 = fmaf(B.z, invN, -m1 * mp);
                        const float c2 = fmaf(B.w, invN, -m2 * mp);
                        ab.x = i00 * c0 + i01 * c1 + i02 * c2;             // :216-218
                        ab.y = i01 * c0 + i11 * c1 + i12 * c2;
                        ab.z = i02 * c0 + i12 * c1 + i22 * c2;
                        ab.w = mp - ab.x * m0 - ab.y * m1 - ab.z * m2;     // :220
                    }
                    float4* slot = ring2 + slot2 * W2 + t;
                    const float4 old = *slot;
                    *slot = ab;
                    acc2 = f4sub(f4add(acc2, ab), old);
                    hbuf2[r * SW + sidx(t + R)] = acc2;  // column sum centred on row v - 2R
                    slot2 = (slot2 + 1 == K) ? 0 : slot2 + 1;
                }
            }
        }
        // (E) epilogue rows from the stage-2 box sums of chunk itn-2
        if (itn >= 2 && colE) {
#pragma unroll
            for (int r = 0; r < CH; r++) {
                const int v = (itn - 2) * CH + r;
                if (v >= 4 * R && v < VH) {
                    const int yq = Y0 + v - 2 * R;
                    const float inv_ny = 1.0f / (float)(min(yq + R, fy1 - 1) - max(yq - R, it.fy) + 1);
                    const float4 S = hout2[r * SW + sidx(t + 2 * R)];
                    const uint32_t g = __ldg(reinterpret_cast<const unsigned int*>(P.guide) + (size_t)yq * P.W + XE);
                    const float s = 1.0f / 255.0f;
                    const float i0 = (float)(g & 0xffu) * s, i1 = (float)((g >> 8) & 0xffu) * s,
                                i2 = (float)((g >> 16) & 0xffu) * s;
                    float q = (S.w + S.x * i0 + S.y * i1 + S.z * i2) * (inv_nxE * inv_ny);  // GuidedFilter.h:243
                    if (P.with_check) {  // StereoEnergy.h:577-610
                        const float xa = __fmul_rn((float)XE, pl.a), yb = __fmul_rn((float)yq, pl.b);
                        float ds = __fadd_rn(__fadd_rn(xa, yb), pl.c);
                        if (!(it.flags & 1)) ds = __fadd_rn(ds, __fmul_rn(0.0f, pl.v));  // channelSum's 4th term
                        const float a5 = __fmul_rn(pl.a, 5.0f), b5 = __fmul_rn(pl.b, 5.0f);
                        const float lo = P.min_disp, hi = P.max_disp;
                        const float dpp = __fadd_rn(__fadd_rn(ds, a5), b5), dpm = __fsub_rn(__fadd_rn(ds, a5), b5);
                        const float dmp = __fadd_rn(__fsub_rn(ds, a5), b5), dmm = __fsub_rn(__fsub_rn(ds, a5), b5);
                        const bool ok = ds >= lo && ds <= hi && dpp >= lo && dpp <= hi && dpm >= lo && dpm <= hi &&
                                        dmp >= lo && dmp <= hi && dmm >= lo && dmm <= hi;
                        if (!ok) q = kCostInvalid;  // CostVolumeEnergy.h:180-182
                    }
                    const int ry = v - 4 * R;  // row inside the tile
                    if (P.out_compact)
                        P.out[(size_t)it.compact_off + (size_t)ry * it.compact_stride + t] = q;
                    else
                        P.out[(size_t)yq * P.out_pitch + XE] = q;
                }
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ H-phase
        for (int task = t; task < Ttot; task += kThreads) {
            const bool st2 = task >= T1;
            const int tk = st2 ? task - T1 : task;
            const int r = tk % CH, k = tk / CH;
            const int chunk = st2 ? itn - 1 : itn;
            if (chunk < 0 || chunk >= nChunks) continue;
            if (!st2 && k >= n1) continue;
            const int v = chunk * CH + r;
            if (v >= VH || v < (st2 ? 4 * R : 2 * R)) continue;
            const float4* in = (st2 ? hbuf2 : hbuf1) + r * SW;
            float4* out = (st2 ? hout2 : hout1) + r * SW;
            const int x0 = (st2 ? 2 * R : R) + k * kRun;  // first output column of the run
            float4 s = in[sidx(x0 - R)];
#pragma unroll 4
            for (int j = 1; j < K; j++) s = f4add(s, in[sidx(x0 - R + j)]);
            out[sidx(x0)] = s;
#pragma unroll
            for (int j = 1; j < kRun; j++) {
                s = f4sub(f4add(s, in[sidx(x0 + R + j)]), in[sidx(x0 - R - 1 + j)]);
                out[sidx(x0 + j)] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K0: one-time guided-filter statistics, GuidedFilter.h:58-102 with T = double.
// Window sums of the 8-bit guide and of its pairwise products are integers: they are summed
// exactly in int32, and only the final normalisation / 3x3 inversion runs in FP64.
// ---------------------------------------------------------------------------------------------
__global__ void lexp_stats_rowsum(const uchar4* __restrict__ guide, int* __restrict__ rs, int H, int W, int R) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int xa = max(x - R, 0), xb = min(x + R, W - 1);
    for (int xx = xa; xx <= xb; xx++) {
        const uchar4 g = guide[(size_t)y * W + xx];
        const int c0 = g.x, c1 = g.y, c2 = g.z;
        s[0] += c0; s[1] += c1; s[2] += c2;
        s[3] += c0 * c0; s[4] += c0 * c1; s[5] += c0 * c2; s[6] += c1 * c1; s[7] += c1 * c2; s[8] += c2 * c2;
    }
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int k = 0; k < 9; k++) rs[k * HW + (size_t)y * W + x] = s[k];
}

__global__ void lexp_stats_finish(const int* __restrict__ rs, float* __restrict__ stats, int H, int W, int R, double eps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t HW = (size_t)H * W;
    long long s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int ya = max(y - R, 0), yb = min(y + R, H - 1);
    for (int yy = ya; yy <= yb; yy++) {
#pragma unroll
        for (int k = 0; k < 9; k++) s[k] += rs[k * HW + (size_t)yy * W + x];
    }
    const double N = (double)((min(x + R, W - 1) - max(x - R, 0) + 1) * (yb - ya + 1));  // :69
    const double sc = 1.0 / 255.0;
    const double m0 = s[0] * sc / N, m1 = s[1] * sc / N, m2 = s[2] * sc / N;             // :70-72
    const double sc2 = sc * sc;
    const double v00 = s[3] * sc2 / N - m0 * m0 + eps, v01 = s[4] * sc2 / N - m0 * m1, v02 = s[5] * sc2 / N - m0 * m2;  // :79-84
    const double v11 = s[6] * sc2 / N - m1 * m1 + eps, v12 = s[7] * sc2 / N - m1 * m2, v22 = s[8] * sc2 / N - m2 * m2 + eps;
    double i00 = v11 * v22 - v12 * v12, i01 = v12 * v02 - v01 * v22, i02 = v01 * v12 - v11 * v02;  // :87-92
    double i11 = v00 * v22 - v02 * v02, i12 = v02 * v01 - v00 * v12, i22 = v00 * v11 - v01 * v01;
    const double det = i00 * v00 + i01 * v01 + i02 * v02;                                          // :94
    const size_t p = (size_t)y * W + x;
    stats[0 * HW + p] = (float)m0; stats[1 * HW + p] = (float)m1; stats[2 * HW + p] = (float)m2;
    stats[3 * HW + p] = (float)(i00 / det); stats[4 * HW + p] = (float)(i01 / det); stats[5 * HW + p] = (float)(i02 / det);
    stats[6 * HW + p] = (float)(i11 / det); stats[7 * HW + p] = (float)(i12 / det); stats[8 * HW + p] = (float)(i22 / det);
}

}  // namespace lexp
