// Batched independent sliding-window problems (BASELINE.json config 4; scripts/ba_all_*.sh run
// many independent tracks).  Placeholder dispatch: each problem goes through the generic
// engine; replaced by the one-CTA-per-window kernel below once it lands.
#include "kernels.cuh"

namespace cslam {

void solve_window_batch(Engine** engines, int n, cslam_summary* summaries) {
    for (int i = 0; i < n; ++i) {
        Engine& e = *engines[i];
        e.upload();
        e.lm_begin();
        e.lm_iterate(e.opt.max_num_iterations + 1, false, summaries ? &summaries[i] : nullptr);
        e.download();
    }
}

}  // namespace cslam
