// TEST INFRASTRUCTURE ONLY.  Self-test of the CUDA emulator (tests/emu/cuda_runtime.h): three toy kernels that use named
// barriers the way lexp_fused_kernel does, one correct and two deliberately broken, to show that the emulator (a) runs a
// correct producer/consumer pipeline, (b) exposes a missing barrier as a schedule-dependent result, (c) reports a wrong
// barrier thread count as a deadlock instead of hanging.
#include <cuda_runtime.h>

namespace {
inline void bar_sync(int id, int n) { emu::barrier(id, n, true); }
inline void bar_arrive(int id, int n) { emu::barrier(id, n, false); }

// 64 threads: warp 0 produces buf[c & 1][lane] = f(c, lane), warp 1 consumes it reversed; double buffered, FULL ids 0/1, EMPTY 2/3
template <int BUG>
__global__ void pipeline(int* out, int nchunks) {
    int* buf = reinterpret_cast<int*>(emu::dyn_smem());  // [2][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = BUG == 2 ? 96 : 64;  // BUG 2: a barrier that expects more threads than ever arrive
    if (warp == 0) {
        for (int c = 0; c < nchunks; c++) {
            if (c >= 2 && BUG != 1) bar_sync(2 + (c & 1), count);  // BUG 1: the producer does not wait for the slot to be drained
            buf[(c & 1) * 32 + lane] = 1000 * c + lane;
            bar_arrive(c & 1, count);
        }
    } else {
        int acc = 0;
        for (int c = 0; c < nchunks; c++) {
            bar_sync(c & 1, count);
            acc += buf[(c & 1) * 32 + (31 - lane)] * (c + 1);
            if (c + 2 < nchunks && BUG != 1) bar_arrive(2 + (c & 1), count);
        }
        out[lane] = acc;
    }
}
}  // namespace

extern "C" {
// returns 0 and fills out[32] on success; 1 and a message on an emulator-detected error
int emu_selftest_run(int bug, int order, int nchunks, int* out, char* msg, int msglen) {
    char ord[8];
    snprintf(ord, sizeof ord, "%d", order);
    setenv("LEXP_EMU_ORDER", ord, 1);
    emu::last_launch_error().clear();
    if (bug == 0) emu::launch(pipeline<0>, dim3(1), dim3(64), 256, out, nchunks);
    else if (bug == 1) emu::launch(pipeline<1>, dim3(1), dim3(64), 256, out, nchunks);
    else emu::launch(pipeline<2>, dim3(1), dim3(64), 256, out, nchunks);
    unsetenv("LEXP_EMU_ORDER");
    if (!emu::last_launch_error().empty()) { snprintf(msg, msglen, "%s", emu::last_launch_error().c_str()); return 1; }
    return 0;
}
}
