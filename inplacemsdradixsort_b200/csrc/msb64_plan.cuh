// msb64_plan.cuh -- the recursion of local_radixsort (msb_64.c:1007-1035) as device-side
// work lists: no host round trip between levels.
//
// After the histogram of level L, one block per segment
//   - scans the segment's bin counts into the write cursors of the scatter kernel
//     (the "device-wide exclusive scan" is per segment and segments are independent,
//     so it is a block scan);
//   - detects a degenerate digit (all keys in one bin, msb_64.c has no such shortcut):
//     the scatter is skipped and the segment moves to the next level where it is;
//   - files every child bucket for the next step: buckets above LOCAL_CAP become
//     segments of level L+1 (with zeroed histograms and a tile list), runs of
//     neighbouring smaller buckets are merged greedily into local-sort units of at
//     most LOCAL_CAP pairs (the role of the reference's size <= 20 / in-cache cut-offs,
//     msb_64.c:1011-1019), and whatever is final but sits in the scratch buffer gets
//     copy tiles.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

constexpr int PLAN_THREADS = 256;
constexpr int PLAN_MAX_BPT = (1 << MAX_BITS) / PLAN_THREADS;

// First kernel of a sort: control block, level-0 segment, its tiles and histogram.
__global__ void init_kernel(const Ctx c, const int bits0)
{
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t gsize = gridDim.x * blockDim.x;
	const uint32_t nt = seg_tile_count(0, c.n);
	if (gtid == 0) {
		Control *ctl = c.ctl;
		for (int l = 0; l <= MAX_LEVELS; ++l) {
			ctl->nsegs[l] = 0;
			ctl->ntiles[l] = 0;
		}
		ctl->nunits = 0;
		ctl->ncopies = 0;
		ctl->error = 0;
		ctl->degenerate = 0;
		ctl->local_pairs = c.n <= LOCAL_CAP ? c.n : 0;
		for (int l = 0; l < MAX_LEVELS; ++l) ctl->moved[l] = 0;
		if (c.n > LOCAL_CAP) {
			ctl->nsegs[0] = 1;
			ctl->ntiles[0] = nt;
			c.segs[0][0] = Seg{0u, c.n, 0u, 0u};
		} else if (c.n > 0) {
			ctl->nunits = 1;
			c.units[0] = Unit{0u, c.n, 0u, 0u};
		}
	}
	if (c.n > LOCAL_CAP) {
		for (uint32_t i = gtid; i < nt; i += gsize) c.tiles[0][i] = Tile{0u, i};
		for (uint32_t i = gtid; i < (1u << bits0); i += gsize) c.hist[0][i] = 0;
	}
}

// bits: digit width of this level; next_bits: of the next level (0 = this is the last).
__global__ void __launch_bounds__(PLAN_THREADS)
plan_kernel(const Ctx c, const int level, const int bits, const int next_bits)
{
	constexpr int THREADS = PLAN_THREADS;
	__shared__ uint32_t s_cnt[1 << MAX_BITS];
	__shared__ uint32_t s_beg[1 << MAX_BITS];
	__shared__ uint32_t s_scratch[THREADS / 32 + 1];
	__shared__ uint32_t s_large[1 << MAX_BITS];     // bins that become segments
	__shared__ uint32_t s_nlarge, s_max, s_seg_base, s_tile_base, s_copy_base;

	const uint32_t tid = threadIdx.x;
	const uint32_t NB = 1u << bits, NBN = next_bits ? (1u << next_bits) : 0u;
	const uint32_t BPT = (NB + THREADS - 1) / THREADS;
	Control *ctl = c.ctl;
	Seg *segs = ((level & 1) ? c.segs[1] : c.segs[0]), *segs_out = ((level & 1) ? c.segs[0] : c.segs[1]);
	Tile *tiles_out = ((level & 1) ? c.tiles[0] : c.tiles[1]);
	uint32_t *hist = ((level & 1) ? c.hist[1] : c.hist[0]), *hist_out = ((level & 1) ? c.hist[0] : c.hist[1]);
	const uint32_t nsegs = ctl->nsegs[level];
	const bool last = next_bits == 0;

	for (uint32_t sg = blockIdx.x; sg < nsegs; sg += gridDim.x) {
		const Seg s = segs[sg];
		uint32_t *h = hist + size_t(sg) * NB;
		if (tid == 0) {
			s_nlarge = 0;
			s_max = 0;
		}
		__syncthreads();

		// counts -> exclusive begins (bins tid*BPT .. tid*BPT+BPT-1)
		uint32_t cnt[PLAN_MAX_BPT], sum = 0, mx = 0;
#pragma unroll
		for (int q = 0; q < PLAN_MAX_BPT; ++q) {
			const uint32_t b = tid * BPT + q;
			cnt[q] = (q < BPT && b < NB) ? h[b] : 0;
			sum += cnt[q];
			mx = max(mx, cnt[q]);
		}
		uint32_t total;
		uint32_t base = block_exclusive_scan<THREADS>(sum, s_scratch, &total);
		atomicMax(&s_max, mx);
		__syncthreads();
		const bool degenerate = s_max == s.size;
		const uint32_t dst_buf = degenerate ? s.buf : (s.buf ^ 1u);

#pragma unroll
		for (int q = 0; q < PLAN_MAX_BPT; ++q) {
			const uint32_t b = tid * BPT + q;
			if (q < BPT && b < NB) {
				const uint32_t beg = s.begin + base;
				s_cnt[b] = cnt[q];
				s_beg[b] = beg;
				h[b] = beg;                       // write cursor of bin b
				base += cnt[q];
				if (!last && cnt[q] > LOCAL_CAP) s_large[atomicAdd(&s_nlarge, 1u)] = b;
			}
		}
		__syncthreads();

		if (tid == 0) {
			if (degenerate) {
				segs[sg].skip = 1;
				atomicAdd(&ctl->degenerate, 1u);
			} else {
				atomicAdd(&ctl->moved[level], s.size);
			}
			if (!last) {
				// greedy merge of neighbouring small buckets into units
				uint32_t run_beg = 0, run_size = 0, local_pairs = 0;
				for (uint32_t b = 0; b <= NB; ++b) {
					const uint32_t cb = b < NB ? s_cnt[b] : 0xffffffffu;
					if (cb == 0) continue;
					if (cb > LOCAL_CAP || run_size + cb > LOCAL_CAP) {
						if (run_size) {
							const uint32_t u = atomicAdd(&ctl->nunits, 1u);
							if (u < c.max_units) c.units[u] = Unit{run_beg, run_size, dst_buf, 0u};
							else atomicOr(&ctl->error, 4u);
							local_pairs += run_size;
						}
						run_size = 0;
						if (cb > LOCAL_CAP) continue;
					}
					if (run_size == 0) run_beg = s_beg[b];
					run_size += cb;
				}
				if (local_pairs) atomicAdd(&ctl->local_pairs, local_pairs);
				// reserve segment slots and tiles for the large children
				uint32_t nt = 0;
				for (uint32_t i = 0; i < s_nlarge; ++i) {
					const uint32_t b = s_large[i];
					nt += seg_tile_count(s_beg[b], s_cnt[b]);
				}
				s_seg_base = atomicAdd(&ctl->nsegs[level + 1], s_nlarge);
				s_tile_base = atomicAdd(&ctl->ntiles[level + 1], nt);
				if (s_seg_base + s_nlarge > c.max_segs) atomicOr(&ctl->error, 1u);
				if (s_tile_base + nt > c.max_tiles) atomicOr(&ctl->error, 2u);
			} else if (dst_buf == 1u) {
				// final data ends in the scratch buffer: copy it home
				const uint32_t nc = (s.size + COPY_TILE - 1) / COPY_TILE;
				s_copy_base = atomicAdd(&ctl->ncopies, nc);
				if (s_copy_base + nc > c.max_copies) atomicOr(&ctl->error, 8u);
			}
		}
		__syncthreads();

		if (!last) {
			const uint32_t nlarge = s_nlarge;
			if (s_seg_base + nlarge <= c.max_segs) {
				uint32_t tile_at = s_tile_base;
				for (uint32_t i = 0; i < nlarge; ++i) {
					const uint32_t b = s_large[i];
					const uint32_t child = s_seg_base + i;
					const uint32_t nt = seg_tile_count(s_beg[b], s_cnt[b]);
					if (tid == 0) segs_out[child] = Seg{s_beg[b], s_cnt[b], dst_buf, 0u};
					for (uint32_t j = tid; j < NBN; j += THREADS)
						hist_out[size_t(child) * NBN + j] = 0;
					if (tile_at + nt <= c.max_tiles)
						for (uint32_t j = tid; j < nt; j += THREADS)
							tiles_out[tile_at + j] = Tile{child, j};
					tile_at += nt;
				}
			}
		} else if (dst_buf == 1u) {
			const uint32_t nc = (s.size + COPY_TILE - 1) / COPY_TILE;
			if (s_copy_base + nc <= c.max_copies)
				for (uint32_t j = tid; j < nc; j += THREADS) {
					const uint32_t off = j * COPY_TILE;
					c.copies[s_copy_base + j] =
						CopyTile{s.begin + off, min(COPY_TILE, s.size - off)};
				}
		}
		__syncthreads();
	}
}

// Buckets that are final but sit in B (all-equal keys after the last digit).
__global__ void __launch_bounds__(256)
copy_kernel(const Ctx c)
{
	const uint32_t ncopies = c.ctl->ncopies;
	for (uint32_t t = blockIdx.x; t < ncopies; t += gridDim.x) {
		const CopyTile ct = c.copies[t];
		for (uint32_t i = threadIdx.x; i < ct.size; i += blockDim.x) {
			st_stream_u64(c.keys[0] + ct.begin + i, ld_stream_u64(c.keys[1] + ct.begin + i));
			st_stream_u64(c.rids[0] + ct.begin + i, ld_stream_u64(c.rids[1] + ct.begin + i));
		}
	}
}

} // namespace msb64
