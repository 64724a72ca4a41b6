// msb64_plan.cuh -- the recursion of local_radixsort (msb_64.c:1007-1035) as device-side
// work lists: no host round trip between levels.
//
// After the histogram of level L, one WARP per segment
//   - scans the segment's bin counts into the write cursors of the scatter kernel
//     (the "device-wide exclusive scan" is per segment and segments are independent,
//     so it is a warp scan over chunks of 32 bins);
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
#include "msb64_local_packed.cuh"   // unit_packable

namespace msb64 {

constexpr int PLAN_THREADS = 256;

// First kernel of a sort: control block, level-0 segment, its tiles and histogram.
__global__ void init_kernel(const Ctx c, const int bits0, const int shift0)
{
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t gsize = gridDim.x * blockDim.x;
	const uint32_t nt = seg_tile_count(c.begin, c.n);
	if (gtid == 0) {
		Control *ctl = c.ctl;
		for (int l = 0; l <= MAX_LEVELS; ++l) {
			ctl->nsegs[l] = 0;
			ctl->ntiles[l] = 0;
			ctl->nready[l] = 0;
		}
		ctl->nunits = 0;
		ctl->nslow = 0;
		ctl->ncopies = 0;
		ctl->error = 0;
		ctl->degenerate = 0;
		ctl->local_pairs = c.n <= LOCAL_CAP ? c.n : 0;
		ctl->hist_keys = c.n > LOCAL_CAP ? c.n : 0;
		for (int l = 0; l < MAX_LEVELS; ++l) ctl->moved[l] = 0;
		if (c.n > LOCAL_CAP) {
			ctl->nsegs[0] = 1;
			ctl->ntiles[0] = nt;
			c.segs[0][0] = Seg{c.begin, c.n, 0u, seg_flags(shift0, 0u)};
			c.segbits[0][0] = SegBits{0ull, ~0ull};
		} else if (c.n > 0) {
			ctl->nslow = 1;                                  // nothing known about the keys: general path
			c.units[c.max_units - 1] = Unit{c.begin, c.n, 0u, 0u};
		}
	}
	if (c.n > LOCAL_CAP) {
		for (uint32_t i = gtid; i < nt; i += gsize) c.tiles[0][i] = Tile{0u, i};
		for (uint32_t i = gtid; i < (1u << bits0); i += gsize) c.hist[0][i] = 0;
		for (uint32_t i = gtid; i < (1u << FUSE_MAX_BITS); i += gsize) c.fused[i] = 0;
	}
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
		if (lane_id() >= d) v += t;
	}
	return v;
}

// Whole warp: make [begin, begin+size) in buffer `buf` a segment of level+1.
// shift: position of the digit the child is partitioned on at the next level;
// ready: the child's histogram (nbn counts) was computed by the fused pass, or NULL.
__device__ __forceinline__ void emit_segment(const Ctx &c, Control *ctl, Seg *segs_out,
					     Tile *tiles_out, uint32_t *hist_out, int level,
					     uint32_t nbn, uint32_t begin, uint32_t size, uint32_t buf,
					     int shift, const uint32_t *ready = nullptr, uint32_t more_flags = 0)
{
	const uint32_t lane = lane_id();
	const uint32_t nt = seg_tile_count(begin, size);
	uint32_t child = 0, tile_at = 0;
	if (lane == 0) {
		child = atomicAdd(&ctl->nsegs[level + 1], 1u);
		tile_at = atomicAdd(&ctl->ntiles[level + 1], nt);
		if (child >= c.max_segs) atomicOr(&ctl->error, 1u);
		else {
			segs_out[child] = Seg{begin, size, buf, seg_flags(shift, (ready ? SEG_HIST_READY : 0u) | more_flags)};
			((level & 1) ? c.segbits[0] : c.segbits[1])[child] = SegBits{0ull, ~0ull};
		}
		if (tile_at + nt > c.max_tiles) atomicOr(&ctl->error, 2u);
		if (ready) atomicAdd(&ctl->nready[level + 1], 1u);
		else atomicAdd(&ctl->hist_keys, (unsigned long long) size);
	}
	child = __shfl_sync(0xffffffffu, child, 0);
	tile_at = __shfl_sync(0xffffffffu, tile_at, 0);
	if (child >= c.max_segs || tile_at + nt > c.max_tiles) return;
	for (uint32_t j = lane; j < nbn; j += 32) hist_out[size_t(child) * nbn + j] = ready ? ready[j] : 0u;
	for (uint32_t j = lane; j < nt; j += 32) tiles_out[tile_at + j] = Tile{child, j};
}

// Whole warp: final data of [begin, begin+size) sits in B, schedule its copy to A.
__device__ __forceinline__ void emit_copy(const Ctx &c, Control *ctl, uint32_t begin, uint32_t size)
{
	const uint32_t lane = lane_id();
	const uint32_t nc = (size + COPY_TILE - 1) / COPY_TILE;
	uint32_t at = 0;
	if (lane == 0) {
		at = atomicAdd(&ctl->ncopies, nc);
		if (at + nc > c.max_copies) atomicOr(&ctl->error, 8u);
	}
	at = __shfl_sync(0xffffffffu, at, 0);
	if (at + nc > c.max_copies) return;
	for (uint32_t j = lane; j < nc; j += 32) {
		const uint32_t off = j * COPY_TILE;
		c.copies[at + j] = CopyTile{begin + off, min(COPY_TILE, size - off)};
	}
}

// bits: digit width of this level; next_bits: of the next level (0 = this is the last);
// fused: c.fused holds the next level's digit counts per bin of this level (level 0 only).
//
// A segment whose digit is degenerate (all keys in one bin) moves to the next level as it is,
// flagged SEG_WANT_BITS: that level's histogram pass then also accumulates OR / AND of its
// keys (c.segbits), and if its digit is degenerate again the plan kernel knows where the
// keys really differ:
//   - nowhere (all keys equal): the segment is finished, nothing below can separate them;
//   - otherwise it goes on with its digit placed right below its highest differing bit,
//     however many bits further down that is.
// Presorted, low-entropy and duplicate-heavy inputs pay two histogram passes for a run of
// dead digits, not one per digit; inputs without degenerate digits pay nothing.
//
// A level that holds a single segment (level 0 always; the first level of every sub-range sort
// of the sharded path) would leave the whole job -- up to 2^bits children, each with a
// histogram row to copy and a tile list to write -- to one warp.  There the warp only claims the
// children's slots and notes them in shared memory; the block's other warps then write the rows
// and tile lists with it (a level-0 plan of 128 children x 2048 tiles: 229 -> ~50 us).
// notes: PLAN_NOTES words of shared memory (the caller's: static in plan_kernel, a corner of the
// dynamic allocation in the tail kernel).
constexpr uint32_t PLAN_NOTES = 3u << MAX_BITS;
__device__ __forceinline__ void plan_pass(const Ctx &c, const int level, const int bits, const int next_bits,
					   const bool fused, uint32_t *notes)
{
	uint32_t *d_child = notes, *d_tile = notes + (1u << MAX_BITS), *d_nt = notes + (2u << MAX_BITS);
	const uint32_t lane = lane_id();
	const uint32_t warps_per_block = PLAN_THREADS / 32;
	const uint32_t NB = 1u << bits, NBN = next_bits ? (1u << next_bits) : 0u;
	Control *ctl = c.ctl;
	Seg *segs = (level & 1) ? c.segs[1] : c.segs[0];
	Seg *segs_out = (level & 1) ? c.segs[0] : c.segs[1];
	Tile *tiles_out = (level & 1) ? c.tiles[0] : c.tiles[1];
	uint32_t *hist = (level & 1) ? c.hist[1] : c.hist[0];
	uint32_t *hist_out = (level & 1) ? c.hist[0] : c.hist[1];
	const SegBits *segbits = (level & 1) ? c.segbits[1] : c.segbits[0];
	const uint32_t nsegs = min(ctl->nsegs[level], c.max_segs);
	const bool single = nsegs == 1;              // the same on every thread of the grid
	if (single) {
		if (blockIdx.x != 0) return;
		for (uint32_t b = threadIdx.x; b < NB; b += PLAN_THREADS) d_child[b] = 0xffffffffu;
		__syncthreads();
	}

	for (uint32_t sg = blockIdx.x * warps_per_block + (threadIdx.x >> 5); sg < nsegs;
	     sg += gridDim.x * warps_per_block) {
		const Seg s = segs[sg];
		uint32_t *h = hist + size_t(sg) * NB;
		const int shift = seg_shift(s.flags);
		// no digit below this one: the keys of a bin are equal, every bucket is final
		const bool last = next_bits == 0 || shift == 0;
		const int child_shift = shift > next_bits ? shift - next_bits : 0;

		// pass 1: is the digit degenerate (one bin holds the whole segment)?
		uint32_t mx = 0;
		for (uint32_t b = lane; b < NB; b += 32) mx = max(mx, h[b]);
		mx = __reduce_max_sync(0xffffffffu, mx);
		if (mx == s.size) {
			if (lane == 0) {
				segs[sg].flags = s.flags | SEG_SKIP;
				atomicAdd(&ctl->degenerate, 1u);
			}
			// bits in which the segment's keys differ (all below `shift`), if this level's
			// histogram pass was asked to collect them
			unsigned long long diff = ~0ull;
			if ((s.flags & SEG_WANT_BITS) && !(s.flags & SEG_HIST_READY)) {
				const SegBits sb = segbits[sg];
				diff = sb.vor & ~sb.vand;
			}
			if (last || diff == 0) {
				if (s.buf == 1u) emit_copy(c, ctl, s.begin, s.size);      // all keys equal: final
				continue;
			}
			// digit of the next level: right below the highest differing bit, or where the
			// schedule puts it when that is not known
			int down = child_shift;
			if (diff != ~0ull) down = min(child_shift, max(63 - __clzll(diff) + 1 - next_bits, 0));
			// the one non-empty bin (for the fused histogram's row, valid at the schedule's position only)
			uint32_t full_bin = 0;
			for (uint32_t b = lane; b < NB; b += 32)
				if (h[b] == s.size) full_bin = b;
			full_bin = __reduce_max_sync(0xffffffffu, full_bin);
			emit_segment(c, ctl, segs_out, tiles_out, hist_out, level, NBN, s.begin, s.size, s.buf, down,
				     fused && down == child_shift ? c.fused + size_t(full_bin) * NBN : nullptr,
				     diff == ~0ull ? SEG_WANT_BITS : 0u);
			continue;
		}
		const uint32_t dst_buf = s.buf ^ 1u;
		if (lane == 0) atomicAdd(&ctl->moved[level], s.size);

		// pass 2: cursors, children
		uint32_t base = s.begin;
		uint32_t run_beg = 0, run_size = 0, run_dig = 0, local_pairs = 0;   // warp-uniform merge state
		// units are collected one per lane and filed 32 at a time with a single atomic
		uint32_t nu = 0, ub = 0, us = 0, uo = 0;
		auto file_units = [&]() {
			if (!nu) return;
			const bool fast = unit_packable(unit_origin(shift, 0u, bits, level == 0));
			uint32_t at = 0;
			if (lane == 0) {
				at = atomicAdd(fast ? &ctl->nunits : &ctl->nslow, nu);
				if (at + nu > c.max_units) atomicOr(&ctl->error, 4u);
			}
			at = __shfl_sync(0xffffffffu, at, 0);
			if (at + nu <= c.max_units && lane < nu)
				c.units[fast ? at + lane : c.max_units - 1 - (at + lane)] = Unit{ub, us, dst_buf, uo};
			nu = 0;
		};
		auto add_unit = [&](uint32_t ubeg, uint32_t usize, uint32_t udig) {
			if (lane == nu) {
				ub = ubeg;
				us = usize;
				uo = unit_origin(shift, udig, bits, level == 0);
			}
			if (++nu == 32) file_units();
		};
		for (uint32_t b0 = 0; b0 < NB; b0 += 32) {
			const uint32_t b = b0 + lane;
			const uint32_t cnt = b < NB ? h[b] : 0;
			const uint32_t inc = warp_inclusive_scan(cnt);
			const uint32_t beg = base + inc - cnt;
			if (b < NB) h[b] = beg;                       // write cursor of bin b
			base += __shfl_sync(0xffffffffu, inc, 31);
			if (last) continue;

			// buckets too large for shared memory: segments of the next level.  One pair of
			// atomics claims segment and tile slots for all of the chunk's children (a
			// round trip per child made the level-0 plan, one warp, the slowest step)
			uint32_t large = __ballot_sync(0xffffffffu, cnt > UNIT_CAP);
			if (large) {
				const bool big = cnt > UNIT_CAP;
				const uint32_t nt_mine = big ? seg_tile_count(beg, cnt) : 0u;
				const uint32_t nt_inc = warp_inclusive_scan(nt_mine);
				const uint32_t nt_all = __shfl_sync(0xffffffffu, nt_inc, 31);
				const uint32_t nchild = __popc(large);
				const uint32_t keys_all = __reduce_add_sync(0xffffffffu, big ? cnt : 0u);
				uint32_t child0 = 0, tile0 = 0;
				if (lane == 0) {
					child0 = atomicAdd(&ctl->nsegs[level + 1], nchild);
					tile0 = atomicAdd(&ctl->ntiles[level + 1], nt_all);
					if (child0 + nchild > c.max_segs) atomicOr(&ctl->error, 1u);
					if (tile0 + nt_all > c.max_tiles) atomicOr(&ctl->error, 2u);
					if (fused) atomicAdd(&ctl->nready[level + 1], nchild);
					else atomicAdd(&ctl->hist_keys, (unsigned long long) keys_all);
				}
				child0 = __shfl_sync(0xffffffffu, child0, 0);
				tile0 = __shfl_sync(0xffffffffu, tile0, 0);
				if (child0 + nchild <= c.max_segs && tile0 + nt_all <= c.max_tiles) {
					const uint32_t my_child = child0 + __popc(large & ((1u << lane) - 1u));
					const uint32_t my_tile = tile0 + nt_inc - nt_mine;
					if (big) {
						segs_out[my_child] = Seg{beg, cnt, dst_buf, seg_flags(child_shift, fused ? SEG_HIST_READY : 0u)};
						((level & 1) ? c.segbits[0] : c.segbits[1])[my_child] = SegBits{0ull, ~0ull};
					}
					if (single) {
						// rows and tile lists: by the whole block, after the loop
						if (big) {
							d_child[b] = my_child;
							d_tile[b] = my_tile;
							d_nt[b] = nt_mine;
						}
						large = 0;
					}
					while (large) {
						const int src = __ffs(large) - 1;
						large &= large - 1;
						const uint32_t child = __shfl_sync(0xffffffffu, my_child, src);
						const uint32_t tile_at = __shfl_sync(0xffffffffu, my_tile, src);
						const uint32_t nt = __shfl_sync(0xffffffffu, nt_mine, src);
						const uint32_t *row = fused ? c.fused + size_t(b0 + src) * NBN : nullptr;
						for (uint32_t j = lane; j < NBN; j += 32) hist_out[size_t(child) * NBN + j] = row ? row[j] : 0u;
						for (uint32_t j = lane; j < nt; j += 32) tiles_out[tile_at + j] = Tile{child, j};
					}
				}
			}
			// greedy merge of neighbouring small buckets into units (all lanes in step)
			uint32_t present = __ballot_sync(0xffffffffu, cnt != 0);
			while (present) {
				const int src = __ffs(present) - 1;
				present &= present - 1;
				const uint32_t cb = __shfl_sync(0xffffffffu, cnt, src);
				const uint32_t bb = __shfl_sync(0xffffffffu, beg, src);
				if (cb > UNIT_CAP || run_size + cb > UNIT_CAP) {
					if (run_size) add_unit(run_beg, run_size, run_dig);
					local_pairs += run_size;
					run_size = 0;
					if (cb > UNIT_CAP) continue;
				}
				if (run_size == 0) {
					run_beg = bb;
					run_dig = b0 + src;
				}
				run_size += cb;
			}
		}
		if (!last) {
			if (run_size) add_unit(run_beg, run_size, run_dig);
			file_units();
			local_pairs += run_size;
			if (local_pairs && lane == 0) atomicAdd(&ctl->local_pairs, local_pairs);
		} else if (dst_buf == 1u) {
			emit_copy(c, ctl, s.begin, s.size);           // final data ends in the scratch buffer
		}
	}
	if (single) {
		__syncthreads();
		for (uint32_t b = threadIdx.x >> 5; b < NB; b += warps_per_block) {
			const uint32_t child = d_child[b];
			if (child == 0xffffffffu) continue;
			const uint32_t tile_at = d_tile[b], nt = d_nt[b];
			const uint32_t *row = fused ? c.fused + size_t(b) * NBN : nullptr;
			for (uint32_t j = lane; j < NBN; j += 32) hist_out[size_t(child) * NBN + j] = row ? row[j] : 0u;
			for (uint32_t j = lane; j < nt; j += 32) tiles_out[tile_at + j] = Tile{child, j};
		}
	}
}

__global__ void __launch_bounds__(PLAN_THREADS)
plan_kernel(const Ctx c, const int level, const int bits, const int next_bits, const bool fused)
{
	__shared__ uint32_t notes[PLAN_NOTES];
	plan_pass(c, level, bits, next_bits, fused, notes);
}

// Buckets that are final but sit in B (all-equal keys after the last digit).
__global__ void __launch_bounds__(256)
copy_kernel(const Ctx c)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		uint32_t e = c.ctl->error;
		if (c.ctl->nunits + c.ctl->nslow > c.max_units) {     // the two unit lists ran into each other
			e |= 4u;
			atomicOr(&c.ctl->error, 4u);
		}
		// last kernel of the sort: tell the host (sticky, read at the device's next use)
		if (e && c.status) *reinterpret_cast<volatile uint32_t *>(c.status) = e;
	}
	const uint32_t ncopies = min(c.ctl->ncopies, c.max_copies);
	for (uint32_t t = blockIdx.x; t < ncopies; t += gridDim.x) {
		const CopyTile ct = c.copies[t];
		for (uint32_t i = threadIdx.x; i < ct.size; i += blockDim.x) {
			st_stream_u64(c.keys[0] + ct.begin + i, ld_stream_u64(c.keys[1] + ct.begin + i));
			st_stream_u64(c.rids[0] + ct.begin + i, ld_stream_u64(c.rids[1] + ct.begin + i));
		}
	}
}

} // namespace msb64
