/* Plain-C restatement of events_to_voxel_grid
 * (RAM_Net/utils/event_tensor_utils.py:71-117).
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg, never by the product.  Pinned against the
 * reference's numpy function run in the build container (tests/golden/voxel_*.npz).
 *
 * Semantics kept from the reference:
 *   - events are [N,4] float64 rows [t, x, y, p]; input is NOT mutated here
 *     (the reference overwrites columns 0 and 3 in place, :95,:100);
 *   - t^ = (B-1)(t - t0)/dT with dT = t[N-1]-t[0], dT==0 -> 1.0 (:88-95), float64;
 *   - ti = (int64) t^  (truncation toward zero, numpy astype(int), :102);
 *   - p==0 -> -1 (:100); left = p(1-dt), right = p*dt in float64;
 *     np.add.at(float32 grid, idx, float64 vals) runs the float64 add loop, so
 *     each step is grid = (float)((double)grid + val), sequentially, all left
 *     votes first (:107-109), then all right votes (:111-113) (probed: casting
 *     val to float first is 1 ulp off on pixels that receive >= 2 votes);
 *   - a vote is dropped iff ti >= B (resp. ti+1 >= B); no lower bound check.
 */
#include <stdint.h>
#include <string.h>

int voxel_oracle(const double *ev, int64_t n, int bins, int width, int height, float *grid)
{
    const int64_t plane = (int64_t)width * height;
    memset(grid, 0, sizeof(float) * (size_t)(plane * bins));
    if (n <= 0) return 0;
    const double t0 = ev[0];
    double dT = ev[4 * (n - 1)] - t0;
    if (dT == 0) dT = 1.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < n; ++i) {
            const double ts = (double)(bins - 1) * (ev[4 * i] - t0) / dT;
            const int64_t x = (int64_t)ev[4 * i + 1];
            const int64_t y = (int64_t)ev[4 * i + 2];
            double p = ev[4 * i + 3];
            if (p == 0) p = -1.0;
            const int64_t ti = (int64_t)ts;
            const double dt = ts - (double)ti;
            if (pass == 0) {
                if (ti < bins) { float *c = &grid[x + y * width + ti * plane]; *c = (float)((double)*c + p * (1.0 - dt)); }
            } else {
                if (ti + 1 < bins) { float *c = &grid[x + y * width + (ti + 1) * plane]; *c = (float)((double)*c + p * dt); }
            }
        }
    }
    return 0;
}
