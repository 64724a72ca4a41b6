// msb64_tail.cuh -- the levels below the schedule's uniform depth in ONE cooperative launch.
//
// The schedule (make_schedule, msb64_b200.cu) gives uniform keys just enough digits to reach
// buckets the local sort finishes; the digits after those ("tail": 7 bits each, down to bit 0)
// are only reached by skewed inputs -- duplicates, shared prefixes, presorted runs.  The host
// cannot know whether a sort needs them (it never synchronises), and launching histogram, plan
// and scatter for every tail level costs ~19 us per EMPTY level: 0.13 ms of a 2^30 sort, but
// 1.8 ms of a sharded step, which runs 16 sub-range sorts of 6 empty tail levels each
// (msb64_shard.cuh).
//
// tail_kernel is a persistent grid (cooperative launch, one grid.sync() between the passes)
// that walks the tail levels with the same device code as the stand-alone kernels
// (plan_pass, scatter_pass; histogram_pass_staged, the histogram pass with its loads staged through
// shared memory, because this grid has too few warps per SM for register loads to fill HBM) and
// leaves at the first level without segments:
// an unused tail costs one launch.  This is the recursion of local_radixsort
// (msb_64.c:1007-1035) continuing below the planned depth.
#pragma once
#include <cooperative_groups.h>

#include "msb64_common.cuh"
#include "msb64_histogram.cuh"
#include "msb64_plan.cuh"
#include "msb64_scatter.cuh"

namespace msb64 {

constexpr int TAIL_BITS = 7;            // width of every tail digit
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_MINB = 3;
static_assert(TAIL_THREADS == PLAN_THREADS, "plan_pass is written for PLAN_THREADS threads");

constexpr size_t tail_smem()
{
	const size_t a = ScatterCfg<TAIL_BITS, TAIL_THREADS>::SMEM, b = HistStagedCfg<TAIL_BITS, TAIL_THREADS>::SMEM,
		     p = size_t(PLAN_NOTES) * 4;
	return a > b ? (a > p ? a : p) : (b > p ? b : p);
}

// levels [first, last]: every one a TAIL_BITS digit; `last` is the schedule's last level (no
// digit below it).
__global__ void __launch_bounds__(TAIL_THREADS, TAIL_MINB)
tail_kernel(const Ctx c, const int first, const int last)
{
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) unsigned char smem_raw[];
	for (int level = first; level <= last; ++level) {
		// written by the plan pass of the level above, final since the grid-wide barrier (or
		// the kernel boundary) behind it: every block takes the same way out
		if (*reinterpret_cast<volatile uint32_t *>(&c.ctl->nsegs[level]) == 0) return;
		histogram_pass_staged<TAIL_BITS, TAIL_THREADS>(c, level);
		grid.sync();
		plan_pass(c, level, TAIL_BITS, level < last ? TAIL_BITS : 0, false, reinterpret_cast<uint32_t *>(smem_raw));
		grid.sync();
		scatter_pass<TAIL_BITS, TAIL_THREADS>(c, level, 0u);
		grid.sync();
	}
}

} // namespace msb64
