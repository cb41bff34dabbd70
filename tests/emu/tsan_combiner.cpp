// TEST INFRASTRUCTURE ONLY (manual tool).  ThreadSanitizer check of the host-side locking of the C-ABI, in particular the
// combiner of concurrent lexp_eval_cell calls: N threads hammer one context with single-cell calls (plus a plan evaluation
// now and then), the library being compiled with -DLEXP_EMU -DLEXP_EMU_NO_KERNELS (kernel launches are no-ops, so the
// results are meaningless; every lock, queue and buffer hand-over is real).
//   g++ -std=c++17 -O1 -g -fsanitize=thread -DLEXP_EMU -DLEXP_EMU_NO_KERNELS -Itests/emu -x c++ localexpstereo_b200/csrc/lexp_capi.cu \
//       tests/emu/tsan_combiner.cpp -o /tmp/tsan_combiner -lpthread && /tmp/tsan_combiner
#include "../../include/lexp_cuda.h"
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>

int main() {
    const int H = 120, W = 160, D = 8, T = 8, ROUNDS = 60;
    lexp_params p{};
    p.height = H; p.width = W; p.ndisp = D; p.windR = 20; p.eps = 1e-4f; p.th_col = 0.5f; p.min_disp = 0; p.max_disp = D - 1;
    lexp_ctx* c = nullptr;
    if (lexp_create(&p, &c)) { std::printf("create failed: %s\n", lexp_last_error()); return 2; }
    std::vector<unsigned char> im((size_t)H * W * 3, 100);
    std::vector<float> vol((size_t)D * H * W, 0.25f);
    lexp_set_image(c, 0, im.data(), W * 3); lexp_set_image(c, 1, im.data(), W * 3);
    lexp_set_volume_host(c, 0, vol.data()); lexp_set_volume_host(c, 1, vol.data());
    std::vector<float> image((size_t)H * W, 0.f);
    std::atomic<int> errors{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int r = 0; r < ROUNDS; r++) {
                // disjoint targets per thread (as the cells of a disjoint group), overlapping filterRects
                lexp_rect targ{10 + 18 * t, 10 + (r % 5) * 20, 15, 15};
                lexp_rect filt{targ.x - 10 < 0 ? 0 : targ.x - 10, targ.y - 10, 35, 35};
                if (filt.x + filt.width > W) filt.width = W - filt.x;
                if (filt.y + filt.height > H) filt.height = H - filt.y;
                lexp_plane pl{0.01f * t, 0.f, 2.f, 0.f};
                float* costs = image.data() + (size_t)filt.y * W + filt.x;
                if (lexp_eval_cell(c, r & 1, &filt, &targ, &pl, costs, W * 4, 1)) errors++;
                if (r % 16 == 5) {  // a batched evaluation from the same threads now and then
                    lexp_plan* plan = nullptr;
                    if (lexp_plan_create(c, 1, &filt, &targ, &plan) == 0) {
                        std::vector<float> tiles(15 * 15);
                        if (lexp_plan_eval_host_tiles(c, plan, 0, &pl, tiles.data(), 1)) errors++;
                        lexp_plan_destroy(plan);
                    }
                }
            }
        });
    for (auto& x : th) x.join();
    int64_t batches = 0, calls = 0;
    lexp_combine_stats(c, &batches, &calls);
    lexp_destroy(c);
    std::printf("tsan_combiner: %d threads x %d rounds, %lld calls combined into %lld launches, errors %d\n", T, ROUNDS, (long long)calls, (long long)batches, errors.load());
    return errors.load() ? 1 : 0;
}
