// msb64_scatter.cuh -- one MSD partition pass (replaces partition_ip / partition_ip_buf,
// msb_64.c:740-978).
//
// The reference permutes in place by cycle following with per-partition cache-line
// buffers; on a GPU that costs bandwidth (every element would be read and written
// through dependent random accesses), so a pass moves the segment from one HBM
// buffer to the other and the last pass / the local sort lands it in the caller's
// arrays.  Algorithmic traffic: 32 bytes per pair (16 read + 16 written).
//
// A block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Per tile of TILE pairs:
//   0. keys and rids of the tile land in shared memory by two bulk asynchronous copies
//      (cp.async.bulk + mbarrier, the non-tensor TMA path) issued by one thread: no
//      register is tied up by a load in flight, and the blocks of an SM (two of 512 threads
//      in the sort's kernels, three of 256 in the tail kernel) cover one another's load
//      latency; tile descriptors are read two tiles ahead;
//   1. rank of every key among the tile's keys with the same digit: one shared-memory
//      atomicAdd on the tile's bin counter (B200 sustains ~9 spread shared atomics
//      per clock per SM, tools/microbench.cu; a ballot multisplit manages ~1); a warp
//      whose digits are all equal takes a single atomic;
//   2. one global atomicAdd per non-empty bin on the segment's write cursor
//      reserves the tile's slice of that bin (the cursors were initialised by the
//      plan kernel from the histogram); a block scan over the bin counts gives the
//      bins' positions in the tile's bin-sorted order.  MSD radix sort is not stable,
//      so the order in which tiles claim their slices is free and no block ever waits
//      for another one;
//   3. the pairs do not move inside shared memory: only a 2-byte source slot per pair
//      is written in bin order (sidx[position] = slot);
//   4. write-out: position i gathers key and rid of slot sidx[i] and stores them to
//      delta[bin] + i, so that consecutive lanes write consecutive addresses.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

template <int BITS, int THREADS>
struct ScatterCfg {
	static constexpr int NB = 1 << BITS;
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr int BPT = (NB + THREADS - 1) / THREADS;   // bins per thread
	static constexpr size_t SMEM = size_t(TILE) * 16               // keys + rids of the tile (bulk-copied)
				       + size_t(TILE) * 2              // source slot of every bin-ordered position
				       + size_t(NB + 32) * 4           // tile-bin counters / bases (+ dummies)
				       + size_t(NB) * 4                // delta
				       + 64 * 4                        // scan scratch
				       + 3 * 32                        // tile descriptors, three deep
				       + 16;                           // two mbarriers (keys, rids)
};

// What a block needs to know about one tile (kept in shared memory, written by thread 0).
struct TileDesc {
	uint32_t seg;     // index of the segment (row of the cursor table)
	uint32_t begin;   // the segment's first element
	uint32_t end;     // one past its last
	uint32_t lo;      // first element slot of the tile (even)
	uint32_t flags;   // bit 0: source buffer, bit 1: segment skipped, bit 2: no such tile
	uint32_t shift;   // position of the segment's digit
	uint32_t pad[2];
};
constexpr uint32_t TD_BUF = 1u, TD_SKIP = 2u, TD_NONE = 4u;

// Rank of a key among the tile's keys with the same digit = old value of the tile's
// bin counter.  The hot loop is branch-free on purpose: ptxas re-materialises the
// shared-window base (S2UR SR_CgaCtaId) in front of every ATOMS that sits behind a
// branch, which made the XU pipe the bottleneck.  A warp whose ITEMS x 32 digits are
// all equal (low-entropy input) claims its ranks with one atomic instead.
// Result per item: rank | digit << 13 (digits NB.. are the dummy bins of invalid items).
constexpr uint32_t RANK_BITS = 13, RANK_MASK = (1u << RANK_BITS) - 1;
template <int ITEMS, int NB>
__device__ __forceinline__ void tile_ranks(uint32_t *cnt, uint32_t (&d)[ITEMS])
{
	const uint32_t d0 = __shfl_sync(0xffffffffu, d[0], 0);
	bool same = true;
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) same = same && d[j] == d0;
	if (__all_sync(0xffffffffu, same)) {
		uint32_t base = 0;
		if (lane_id() == 0) base = atomicAdd(&cnt[d0], uint32_t(32 * ITEMS));
		base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) d[j] = (base + j * 32 + lane_id()) | (d0 << RANK_BITS);
	} else {
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) d[j] = atomicAdd(&cnt[d[j]], 1u) | (d[j] << RANK_BITS);
	}
}

// (the pass as a device function: scatter_kernel below and the cooperative tail kernel,
// msb64_tail.cuh, run the same code)
template <int BITS, int THREADS>
__device__ __forceinline__ void scatter_pass(const Ctx &c, const int level, const uint32_t origin)
{
	using Cfg = ScatterCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, ITEMS = Cfg::ITEMS, BPT = Cfg::BPT;
	static_assert(ITEMS % 2 == 0, "keys are read from shared memory as 16-byte pairs");
	static_assert(TILE <= RANK_MASK + 1 && TILE <= 65536, "rank / slot packing");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *kin = reinterpret_cast<uint64_t *>(smem_raw);         // [TILE] keys of the tile
	uint64_t *rin = kin + TILE;                                     // [TILE] rids of the tile
	uint16_t *sidx = reinterpret_cast<uint16_t *>(rin + TILE);      // [TILE] source slot by bin-ordered position
	uint32_t *cnt = reinterpret_cast<uint32_t *>(sidx + TILE);      // [NB + 32] counts, then local bases
	uint32_t *delta = cnt + NB + 32;                                // [NB] global - local base
	uint32_t *scratch = delta + NB;                                 // [64]
	TileDesc *sdesc = reinterpret_cast<TileDesc *>(scratch + 64);   // [3]
	uint64_t *bar = reinterpret_cast<uint64_t *>(sdesc + 3);

	const uint32_t tid = threadIdx.x;
	const uint32_t ntiles = c.ctl->ntiles[level];
	const Seg *segs = (level & 1) ? c.segs[1] : c.segs[0];
	const Tile *tiles = (level & 1) ? c.tiles[1] : c.tiles[0];
	uint32_t *cursors = (level & 1) ? c.hist[1] : c.hist[0];
	const uint32_t G = gridDim.x;
	if (blockIdx.x >= ntiles) return;
	// every segment of the level kept its place (degenerate digits: presorted keys, shared
	// prefixes): nothing to walk through tile by tile
	if (*reinterpret_cast<volatile uint32_t *>(&c.ctl->moved[level]) == 0) return;

	// thread 0 only: descriptor of tile t into slot
	auto fetch_desc = [&](uint32_t t, TileDesc *slot) {
		if (t >= ntiles) {
			slot->flags = TD_NONE;
			return;
		}
		const Tile tile = tiles[t];
		const Seg s = segs[tile.seg];
		slot->seg = tile.seg;
		slot->begin = s.begin;
		slot->end = s.begin + s.size;
		slot->lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		slot->flags = (s.buf ? TD_BUF : 0u) | ((s.flags & SEG_SKIP) ? TD_SKIP : 0u);
		slot->shift = uint32_t(seg_shift(s.flags));
	};
	// Slots of a tile's window that bulk copies fetch: up to the segment's end (rounded to the
	// 16-byte granule), so that the short last tile of a segment does not cost a whole tile
	// of traffic.  0 = the window would leave the array (last tile only): plain loads instead.
	auto window = [&](uint32_t lo, uint32_t end) -> uint32_t {
		const uint32_t w = min(uint32_t(TILE), (end - lo + 1u) & ~1u);
		return lo + w <= c.end ? w : 0u;
	};
	// thread 0 only (slots in front of the segment receive the neighbours' data and are ignored).
	// Keys and rids travel separately, each behind its own mbarrier: the keys' buffer is free
	// as soon as the keys have been written out, the rids are not needed before the write-out.
	auto start_keys = [&](const TileDesc &d) {
		if (d.flags & (TD_SKIP | TD_NONE)) return;
		const uint32_t w = window(d.lo, d.end);
		if (!w) return;
		mbar_expect_tx(&bar[0], w * 8);
		bulk_copy_g2s(kin, ((d.flags & TD_BUF) ? c.keys[1] : c.keys[0]) + d.lo, w * 8, &bar[0]);
	};
	auto start_rids = [&](const TileDesc &d) {
		if (d.flags & (TD_SKIP | TD_NONE)) return;
		const uint32_t w = window(d.lo, d.end);
		if (!w) return;
		mbar_expect_tx(&bar[1], w * 8);
		bulk_copy_g2s(rin, ((d.flags & TD_BUF) ? c.rids[1] : c.rids[0]) + d.lo, w * 8, &bar[1]);
	};

	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		fetch_desc(blockIdx.x, &sdesc[0]);
		fetch_desc(blockIdx.x + G, &sdesc[1]);
		start_keys(sdesc[0]);
		start_rids(sdesc[0]);
	}
	for (int i = tid; i < NB + 32; i += THREADS) cnt[i] = 0;
	__syncthreads();

	uint32_t parity_k = 0, parity_r = 0, slot = 0;
	for (uint32_t t = blockIdx.x; t < ntiles; t += G, slot = slot == 2 ? 0 : slot + 1) {
		const TileDesc cur = sdesc[slot];
		TileDesc *next_slot = &sdesc[slot == 2 ? 0 : slot + 1];
		TileDesc *after_slot = &sdesc[slot == 0 ? 2 : slot - 1];       // slot + 2 mod 3
		if (cur.flags & TD_SKIP) {
			if (tid == 0) {
				start_keys(*next_slot);
				start_rids(*next_slot);
				fetch_desc(t + 2 * G, after_slot);
			}
			__syncthreads();
			continue;
		}
		const bool src_b = cur.flags & TD_BUF;
		uint64_t *dst_keys = src_b ? c.keys[0] : c.keys[1];
		uint64_t *dst_rids = src_b ? c.rids[0] : c.rids[1];
		const uint32_t lo = cur.lo;
		const int shift = int(cur.shift);
		const bool full = lo >= cur.begin && lo + TILE <= cur.end;
		const uint32_t count = min(lo + TILE, cur.end) - max(lo, cur.begin);

		// 0. the tile's keys in shared memory (the rids are awaited in front of their write-out)
		const bool bulk = window(lo, cur.end) != 0;
		if (bulk) {
			mbar_wait(&bar[0], parity_k);
			parity_k ^= 1u;
		} else {
			// the window would cross the end of the array (last tile only): plain loads
			const uint64_t *src_keys = src_b ? c.keys[1] : c.keys[0];
			const uint64_t *src_rids = src_b ? c.rids[1] : c.rids[0];
			for (uint32_t i = tid; i < TILE; i += THREADS) {
				const uint32_t e = lo + i;
				if (e >= cur.begin && e < cur.end) {
					kin[i] = ld_stream_u64(src_keys + e);
					rin[i] = ld_stream_u64(src_rids + e);
				}
			}
			__syncthreads();
		}

		// 1. digits, ranks inside the tile's bins.  Thread `tid` owns slots
		//    (jj * THREADS + tid) * 2 + {0, 1}; items outside the segment count into a
		//    per-lane dummy bin behind the real ones
		uint32_t dr[ITEMS];
		{
			const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(kin);
#pragma unroll
			for (int jj = 0; jj < ITEMS / 2; ++jj) {
				const ulonglong2 v = k2[jj * THREADS + tid];
				dr[2 * jj] = (uint32_t(v.x >> shift) - origin) & (NB - 1);
				dr[2 * jj + 1] = (uint32_t(v.y >> shift) - origin) & (NB - 1);
			}
			if (!full) {
#pragma unroll
				for (int j = 0; j < ITEMS; ++j) {
					const uint32_t e = lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
					if (e < cur.begin || e >= cur.end) dr[j] = NB + lane_id();
				}
			}
		}
		tile_ranks<ITEMS, NB>(cnt, dr);
		__syncthreads();

		// 2. per bin: claim the tile's slice of the segment's bin, exclusive scan over bins
		uint32_t tot[BPT], g[BPT], sum = 0;
#pragma unroll
		for (int q = 0; q < BPT; ++q) {
			const int b = tid * BPT + q;
			tot[q] = b < NB ? cnt[b] : 0;
			g[q] = tot[q] ? atomicAdd(&cursors[size_t(cur.seg) * NB + b], tot[q]) : 0;
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

		// 3. bin-ordered position -> source slot
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t d = dr[j] >> RANK_BITS;
			if (d < NB) sidx[cnt[d] + (dr[j] & RANK_MASK)] = uint16_t(((j >> 1) * THREADS + tid) * 2 + (j & 1));
		}
		__syncthreads();

		// 4. coalesced write-out: position i of the bin-ordered tile goes to delta[bin] + i.
		//    Keys first; once every thread has read its keys the keys of the NEXT tile are
		//    requested into the same buffer (they land while the rids go out and the
		//    counters are cleared), then the rids, then the next tile's rids are requested
		//    (they land while the next tile is being ranked).
		uint32_t dst[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < count) {
				const uint64_t key = kin[sidx[i]];
				dst[j] = delta[(uint32_t(key >> shift) - origin) & (NB - 1)] + i;
				st_stream_u64(dst_keys + dst[j], key);
			}
		}
		__syncthreads();
		if (tid == 0) start_keys(*next_slot);
		if (bulk) {
			mbar_wait(&bar[1], parity_r);
			parity_r ^= 1u;
		}
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < count) st_stream_u64(dst_rids + dst[j], rin[sidx[i]]);
		}
		// meanwhile: counters back to zero for the next tile, descriptor of the tile after it
		for (int i = tid; i < NB + 32; i += THREADS) cnt[i] = 0;
		if (tid == 0) fetch_desc(t + 2 * G, after_slot);
		__syncthreads();
		if (tid == 0) start_rids(*next_slot);
	}
	// nothing is in flight (a tile's copies are awaited before the next ones are requested):
	// give the barriers' words back, the tail kernel runs other passes in this memory
	if (tid == 0) {
		mbar_inval(&bar[0]);
		mbar_inval(&bar[1]);
	}
}

template <int BITS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
scatter_kernel(const Ctx c, const int level, const uint32_t origin)
{
	scatter_pass<BITS, THREADS>(c, level, origin);
}

} // namespace msb64
