// msb64_scatter.cuh -- one MSD partition pass (replaces partition_ip / partition_ip_buf,
// msb_64.c:740-978).
//
// The reference permutes in place by cycle following with per-partition cache-line
// buffers; on a GPU that costs bandwidth (every element would be read and written
// through dependent random accesses), so a pass moves the segment from one HBM
// buffer to the other and the last pass / the local sort lands it in the caller's
// arrays.  Algorithmic traffic: 32 bytes per pair (16 read + 16 written).
//
// Per tile of TILE pairs:
//   1. 16-byte coalesced loads of the keys into registers;
//   2. rank of every key among the tile's keys with the same digit: ballots find
//      the digit peers inside the warp, per-warp counters in shared memory (no
//      atomics) count across the thread's items, a scan over warps and bins turns
//      them into positions in the tile's bin-sorted order;
//   3. one global atomicAdd per non-empty bin on the segment's write cursor
//      reserves the tile's slice of that bin (the cursors were initialised by the
//      plan kernel from the histogram).  MSD radix sort is not stable, so the
//      order in which tiles claim their slices is free and no block ever waits
//      for another one;
//   4. keys and rids are staged through shared memory in bin order and written
//      out so that consecutive lanes write consecutive addresses.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

template <int BITS, int THREADS>
struct ScatterCfg {
	static constexpr int NB = 1 << BITS;
	static constexpr int WARPS = THREADS / 32;
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr int BPT = (NB + THREADS - 1) / THREADS;   // bins per thread
	static constexpr size_t SMEM = size_t(TILE) * 16               // staged keys + rids
				       + size_t(WARPS) * NB * 4        // per-warp counters
				       + size_t(NB) * 4                // delta
				       + 64 * 4;                       // scan scratch
};

template <int BITS, int THREADS>
__global__ void __launch_bounds__(THREADS)
scatter_kernel(const Ctx c, const int level, const int shift)
{
	using Cfg = ScatterCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, WARPS = Cfg::WARPS, ITEMS = Cfg::ITEMS, BPT = Cfg::BPT;
	static_assert(ITEMS % 2 == 0, "tile is loaded as 16-byte pairs");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);       // [TILE]
	uint64_t *srids = skeys + TILE;                                 // [TILE]
	uint32_t *wcnt = reinterpret_cast<uint32_t *>(srids + TILE);    // [WARPS][NB]
	uint32_t *delta = wcnt + WARPS * NB;                            // [NB]
	uint32_t *scratch = delta + NB;                                 // [64]

	const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
	const uint32_t ntiles = c.ctl->ntiles[level];
	const Seg *segs = ((level & 1) ? c.segs[1] : c.segs[0]);
	const Tile *tiles = ((level & 1) ? c.tiles[1] : c.tiles[0]);
	uint32_t *cursors = ((level & 1) ? c.hist[1] : c.hist[0]);
	uint32_t *mine = wcnt + warp * NB;
	const uint32_t lt = lanemask_lt();

	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const Tile tile = tiles[t];
		const Seg s = segs[tile.seg];
		if (s.skip) continue;
		const uint64_t *src_keys = (s.buf ? c.keys[1] : c.keys[0]), *src_rids = (s.buf ? c.rids[1] : c.rids[0]);
		uint64_t *dst_keys = (s.buf ? c.keys[0] : c.keys[1]), *dst_rids = (s.buf ? c.rids[0] : c.rids[1]);
		const uint32_t end = s.begin + s.size;
		const uint32_t lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		const bool full = lo >= s.begin && lo + TILE <= end;
		const uint32_t vlo = max(lo, s.begin), vhi = min(lo + TILE, end);
		const uint32_t count = vhi - vlo;

		for (int i = tid; i < WARPS * NB; i += THREADS) wcnt[i] = 0;
		__syncthreads();

		// 1. keys
		uint64_t k[ITEMS];
		uint32_t validmask = 0;      // bit j: item j is inside the segment
		if (full) {
#pragma unroll
			for (int j = 0; j < ITEMS / 2; ++j) {
				const ulonglong2 v = ld_stream_u64x2(src_keys + lo + (j * THREADS + tid) * 2);
				k[2 * j] = v.x;
				k[2 * j + 1] = v.y;
			}
			validmask = (1u << ITEMS) - 1;
		} else {
#pragma unroll
			for (int j = 0; j < ITEMS; ++j) {
				const uint32_t e = lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
				const bool valid = e >= s.begin && e < end;
				k[j] = valid ? ld_stream_u64(src_keys + e) : 0;
				validmask |= uint32_t(valid) << j;
			}
		}

		// 2. rank inside the warp, per-warp running counts across items
		uint32_t rank[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t d = uint32_t(k[j] >> shift) & (NB - 1);
			const bool valid = (validmask >> j) & 1u;
			uint32_t peers = match_digit<BITS>(d);
			if (!full) peers &= __ballot_sync(0xffffffffu, valid);
			const uint32_t leader = __ffs(peers) - 1;
			uint32_t before = 0;
			if (valid && lane == leader) {
				before = mine[d];
				mine[d] = before + __popc(peers);
			}
			before = __shfl_sync(0xffffffffu, before, leader & 31);
			rank[j] = before + __popc(peers & lt);
			__syncwarp();
		}
		__syncthreads();

		// 3. per bin: exclusive scan over warps, over bins, global slice
		uint32_t tot[BPT], sum = 0;
#pragma unroll
		for (int q = 0; q < BPT; ++q) {
			const int b = tid * BPT + q;
			uint32_t run = 0;
			if (b < NB) {
#pragma unroll
				for (int w = 0; w < WARPS; ++w) {
					const uint32_t v = wcnt[w * NB + b];
					wcnt[w * NB + b] = run;
					run += v;
				}
			}
			tot[q] = run;
			sum += run;
		}
		uint32_t total;
		uint32_t lbase = block_exclusive_scan<THREADS>(sum, scratch, &total);
#pragma unroll
		for (int q = 0; q < BPT; ++q) {
			const int b = tid * BPT + q;
			if (b < NB) {
				uint32_t g = 0;
				if (tot[q]) g = atomicAdd(&cursors[size_t(tile.seg) * NB + b], tot[q]);
				delta[b] = g - lbase;
#pragma unroll
				for (int w = 0; w < WARPS; ++w) wcnt[w * NB + b] += lbase;
				lbase += tot[q];
			}
		}
		__syncthreads();

		// 4a. stage keys in bin order, fetch rids meanwhile
		uint32_t pos[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t d = uint32_t(k[j] >> shift) & (NB - 1);
			pos[j] = mine[d] + rank[j];
		}
		uint64_t r[ITEMS];
		if (full) {
#pragma unroll
			for (int j = 0; j < ITEMS / 2; ++j) {
				const ulonglong2 v = ld_stream_u64x2(src_rids + lo + (j * THREADS + tid) * 2);
				r[2 * j] = v.x;
				r[2 * j + 1] = v.y;
			}
		} else {
#pragma unroll
			for (int j = 0; j < ITEMS; ++j) {
				const uint32_t e = lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
				r[j] = ((validmask >> j) & 1u) ? ld_stream_u64(src_rids + e) : 0;
			}
		}
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if ((validmask >> j) & 1u) skeys[pos[j]] = k[j];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if ((validmask >> j) & 1u) srids[pos[j]] = r[j];
		__syncthreads();

		// 4b. coalesced write-out: slot i of the staged tile goes to delta[bin] + i
		for (uint32_t i = tid; i < count; i += THREADS) {
			const uint64_t key = skeys[i];
			const uint32_t d = uint32_t(key >> shift) & (NB - 1);
			const uint32_t dst = delta[d] + i;
			st_stream_u64(dst_keys + dst, key);
			st_stream_u64(dst_rids + dst, srids[i]);
		}
		__syncthreads();
	}
}

} // namespace msb64
