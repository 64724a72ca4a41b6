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
//   2. rank of every key among the tile's keys with the same digit: one shared-memory
//      atomicAdd on the tile's bin counter (B200 sustains ~9 spread shared atomics
//      per clock per SM, tools/microbench.cu; a ballot multisplit manages ~1); a warp
//      whose digits are all equal takes a single atomic; a block scan over the
//      bin counts turns the ranks into positions in the tile's bin-sorted order;
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
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr int BPT = (NB + THREADS - 1) / THREADS;   // bins per thread
	static constexpr size_t SMEM = size_t(TILE) * 16               // staged keys + rids
				       + size_t(NB + 32) * 4           // tile-bin counters / bases (+ dummies)
				       + size_t(NB) * 4                // delta
				       + 64 * 4;                       // scan scratch
};

// Rank of a key among the tile's keys with the same digit = old value of the tile's
// bin counter.  The hot loop is branch-free on purpose: ptxas re-materialises the
// shared-window base (S2UR SR_CgaCtaId) in front of every ATOMS that sits behind a
// branch, which made the XU pipe the bottleneck.  A warp whose ITEMS x 32 digits are
// all equal (low-entropy input) claims its ranks with one atomic instead.
template <int ITEMS, int NB>
__device__ __forceinline__ void tile_ranks(uint32_t *cnt, const uint64_t (&k)[ITEMS], int shift,
					   uint32_t validmask, uint32_t (&rank)[ITEMS])
{
	// items outside the segment count into a per-lane dummy bin behind the real ones
	uint32_t d[ITEMS];
#pragma unroll
	for (int j = 0; j < ITEMS; ++j)
		d[j] = ((validmask >> j) & 1u) ? (uint32_t(k[j] >> shift) & (NB - 1)) : NB + lane_id();
	const uint32_t d0 = __shfl_sync(0xffffffffu, d[0], 0);
	bool same = true;
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) same = same && d[j] == d0;
	if (__all_sync(0xffffffffu, same)) {
		uint32_t base = 0;
		if (lane_id() == 0) base = atomicAdd(&cnt[d0], uint32_t(32 * ITEMS));
		base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) rank[j] = base + j * 32 + lane_id();
	} else {
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) rank[j] = atomicAdd(&cnt[d[j]], 1u);
	}
}

template <int BITS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
scatter_kernel(const Ctx c, const int level, const int shift)
{
	using Cfg = ScatterCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, ITEMS = Cfg::ITEMS, BPT = Cfg::BPT;
	static_assert(ITEMS % 2 == 0, "tile is loaded as 16-byte pairs");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);       // [TILE]
	uint64_t *srids = skeys + TILE;                                 // [TILE]
	uint32_t *cnt = reinterpret_cast<uint32_t *>(srids + TILE);     // [NB] counts, then local bases
	uint32_t *delta = cnt + NB + 32;                                // [NB] global - local base
	uint32_t *scratch = delta + NB;                                 // [64]

	const uint32_t tid = threadIdx.x;
	const uint32_t ntiles = c.ctl->ntiles[level];
	const Seg *segs = (level & 1) ? c.segs[1] : c.segs[0];
	const Tile *tiles = (level & 1) ? c.tiles[1] : c.tiles[0];
	uint32_t *cursors = (level & 1) ? c.hist[1] : c.hist[0];

	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const Tile tile = tiles[t];
		const Seg s = segs[tile.seg];
		if (s.skip) continue;
		const uint64_t *src_keys = s.buf ? c.keys[1] : c.keys[0];
		const uint64_t *src_rids = s.buf ? c.rids[1] : c.rids[0];
		uint64_t *dst_keys = s.buf ? c.keys[0] : c.keys[1];
		uint64_t *dst_rids = s.buf ? c.rids[0] : c.rids[1];
		const uint32_t end = s.begin + s.size;
		const uint32_t lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		const bool full = lo >= s.begin && lo + TILE <= end;
		const uint32_t count = min(lo + TILE, end) - max(lo, s.begin);

		for (int i = tid; i < NB + 32; i += THREADS) cnt[i] = 0;
		__syncthreads();

		// 1. keys -> registers, 2. rank inside the tile's bin
		uint64_t k[ITEMS];
		uint32_t rank[ITEMS];
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
		tile_ranks<ITEMS, NB>(cnt, k, shift, validmask, rank);
		// rids: issued now, consumed after the scan
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
		__syncthreads();

		// 3. per bin: claim the tile's slice of the segment's bin, exclusive scan over bins
		uint32_t tot[BPT], g[BPT], sum = 0;
#pragma unroll
		for (int q = 0; q < BPT; ++q) {
			const int b = tid * BPT + q;
			tot[q] = b < NB ? cnt[b] : 0;
			g[q] = tot[q] ? atomicAdd(&cursors[size_t(tile.seg) * NB + b], tot[q]) : 0;
			sum += tot[q];
		}
		uint32_t total;
		uint32_t lbase = block_exclusive_scan<THREADS>(sum, scratch, &total);
#pragma unroll
		for (int q = 0; q < BPT; ++q) {
			const int b = tid * BPT + q;
			if (b < NB) {
				cnt[b] = lbase;
				delta[b] = g[q] - lbase;
				lbase += tot[q];
			}
		}
		__syncthreads();

		// 4a. stage keys and rids in bin order
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if ((validmask >> j) & 1u) {
				const uint32_t p = cnt[uint32_t(k[j] >> shift) & (NB - 1)] + rank[j];
				skeys[p] = k[j];
				srids[p] = r[j];
			}
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
