// msb64_histogram.cuh -- per-segment digit histogram (replaces histogram(), msb_64.c:701-738).
//
// Algorithmic traffic: 8 bytes read per key, nothing written but the counters.
// One block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; it keeps one NB-bin
// histogram in shared memory, updated with shared-memory atomics (measured on B200:
// ~9 spread atomics per clock per SM, tools/microbench.cu -- far above the 2.3 keys
// per clock per SM that HBM can deliver, and 9x the rate of a ballot multisplit).
// A warp whose 32 keys all carry the same digit (presorted or low-entropy inputs)
// adds once instead of serialising 32 atomics on one address.  The block histogram is
// added to the segment's global counters whenever the block moves to another segment.
// `origin`: digit of the smallest possible key at this level (level 0 of a sort whose key
// range is known, see make_schedule in msb64_b200.cu; 0 everywhere else): the digit is
// (key >> shift) - origin.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

template <int BITS, int THREADS>
struct HistCfg {
	static constexpr int NB = 1 << BITS;
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr size_t SMEM = size_t(NB + 32 + 4) * sizeof(uint32_t);   // + per-lane dummy bins + OR / AND words
};

// The tile's digits of one thread into the shared histogram.  Branch-free hot loop (a
// branch in front of a shared atomic makes ptxas re-materialise the shared-window base,
// S2UR SR_CgaCtaId, per atomic); a warp whose ITEMS x 32 digits are all equal (presorted
// or low-entropy input) adds once instead of serialising on one address.
// FUSE: also count the next level's digit per bin of this level (table h2, [NB][2^fbits]).
template <int ITEMS, int NB, bool FUSE>
__device__ __forceinline__ void hist_add_tile(uint32_t *h, const uint64_t (&k)[ITEMS], int shift,
					      uint32_t origin, uint32_t validmask, uint32_t *h2 = nullptr,
					      int fshift = 0, int fbits = 0)
{
	// keys outside the segment count into a per-lane dummy bin behind the real ones
	uint32_t d[ITEMS];
#pragma unroll
	for (int j = 0; j < ITEMS; ++j)
		d[j] = ((validmask >> j) & 1u) ? ((uint32_t(k[j] >> shift) - origin) & (NB - 1)) : NB + lane_id();
	const uint32_t d0 = __shfl_sync(0xffffffffu, d[0], 0);
	bool same = true;
#pragma unroll
	for (int j = 0; j < ITEMS; ++j) same = same && d[j] == d0;
	if (__all_sync(0xffffffffu, same)) {
		if (lane_id() == 0) atomicAdd(&h[d0], uint32_t(32 * ITEMS));
	} else {
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) atomicAdd(&h[d[j]], 1u);
	}
	if (FUSE) {
		// invalid items (digit >= NB) land in the dummy rows behind the table
		const uint32_t fmask = (1u << fbits) - 1;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			atomicAdd(&h2[(d[j] << fbits) | (uint32_t(k[j] >> fshift) & fmask)], 1u);
	}
}

// For segments flagged SEG_WANT_BITS the pass also accumulates OR and AND of the keys
// (c.segbits): the plan kernel learns from them where the segment's keys really differ.
// FUSE (level 0 only, one segment): fbits = width of the level-1 digit (right below this level's); the counts of
// level-1 digits per level-0 bin go to c.fused and become the children's histograms in the
// plan kernel, so level 1 needs no histogram pass of its own.
// (the pass as a device function: histogram_kernel below and the cooperative tail kernel,
// msb64_tail.cuh, run the same code)
template <int BITS, int THREADS, bool FUSE>
__device__ __forceinline__ void histogram_pass(const Ctx &c, const int level, const uint32_t origin, const int fbits)
{
	using Cfg = HistCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, ITEMS = Cfg::ITEMS;
	static_assert(ITEMS % 2 == 0, "tile is loaded as 16-byte pairs");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint32_t *sh = reinterpret_cast<uint32_t *>(smem_raw);     // [NB + 32]
	uint32_t *sbits = sh + NB + 32;                            // [4] OR lo, OR hi, AND lo, AND hi of the block's keys
	uint32_t *sh2 = sbits + 4;                                 // FUSE: [(NB + 32) << fbits] (dummy rows included)

	const uint32_t tid = threadIdx.x;
	const uint32_t ntiles = c.ctl->ntiles[level];
	if (!FUSE && c.ctl->nready[level] == c.ctl->nsegs[level]) return;    // every histogram came from the fused pass
	const Seg *segs = (level & 1) ? c.segs[1] : c.segs[0];
	const Tile *tiles = (level & 1) ? c.tiles[1] : c.tiles[0];
	uint32_t *hist = (level & 1) ? c.hist[1] : c.hist[0];
	SegBits *segbits = (level & 1) ? c.segbits[1] : c.segbits[0];
	// OR / AND of the keys this thread has seen of the current segment
	unsigned long long bor = 0, band = ~0ull;
	// block histogram and OR / AND words to the segment's global counters (all threads)
	bool want = false;            // the current segment asked for OR / AND
	auto flush = [&](uint32_t seg, bool again) {
		if (want) {
		const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(bor));
		const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(bor >> 32));
		const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(band));
		const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(band >> 32));
		if (lane_id() == 0) {
			atomicOr(&sbits[0], olo);
			atomicOr(&sbits[1], ohi);
			atomicAnd(&sbits[2], alo);
			atomicAnd(&sbits[3], ahi);
		}
		bor = 0;
		band = ~0ull;
		}
		__syncthreads();
		for (int b = tid; b < NB; b += THREADS) {
			const uint32_t v = sh[b];
			if (v) {
				atomicAdd(&hist[size_t(seg) * NB + b], v);
				sh[b] = 0;
			}
		}
		if (want && tid == 0) {
			atomicOr(&segbits[seg].vor, ((unsigned long long) sbits[1] << 32) | sbits[0]);
			atomicAnd(&segbits[seg].vand, ((unsigned long long) sbits[3] << 32) | sbits[2]);
			sbits[0] = sbits[1] = 0u;
			sbits[2] = sbits[3] = 0xffffffffu;
		}
		if (again) __syncthreads();
	};

	for (int i = tid; i < NB + 32; i += THREADS) sh[i] = 0;
	if (tid < 4) sbits[tid] = tid < 2 ? 0u : 0xffffffffu;
	if (FUSE)
		for (int i = tid; i < ((NB + 32) << fbits); i += THREADS) sh2[i] = 0;
	__syncthreads();

	uint32_t cur_seg = 0xffffffffu;
	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const Tile tile = tiles[t];
		if (tile.seg != cur_seg) {
			if (cur_seg != 0xffffffffu) flush(cur_seg, true);
			cur_seg = tile.seg;
		}
		const Seg s = segs[tile.seg];
		want = (s.flags & (SEG_WANT_BITS | SEG_HIST_READY)) == SEG_WANT_BITS;
		if (s.flags & SEG_HIST_READY) continue;            // counted by the fused pass of the level above
		const int shift = seg_shift(s.flags);
		const uint64_t *keys = s.buf ? c.keys[1] : c.keys[0];
		const uint32_t end = s.begin + s.size;
		const uint32_t lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		const bool full = lo >= s.begin && lo + TILE <= end;

		uint64_t k[ITEMS];
		uint32_t validmask = 0;
		if (full) {
#pragma unroll
			for (int j = 0; j < ITEMS / 2; ++j) {
				const ulonglong2 v = ld_stream_u64x2(keys + lo + (j * THREADS + tid) * 2);
				k[2 * j] = v.x;
				k[2 * j + 1] = v.y;
			}
			validmask = (1u << ITEMS) - 1;
		} else {
#pragma unroll
			for (int j = 0; j < ITEMS; ++j) {
				const uint32_t e = lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
				const bool valid = e >= s.begin && e < end;
				k[j] = valid ? ld_stream_u64(keys + e) : 0;
				validmask |= uint32_t(valid) << j;
			}
		}
		if (want) {
#pragma unroll
			for (int j = 0; j < ITEMS; ++j)
				if ((validmask >> j) & 1u) {
					bor |= k[j];
					band &= k[j];
				}
		}
		hist_add_tile<ITEMS, NB, FUSE>(sh, k, shift, origin, validmask, sh2, shift - fbits, fbits);
	}
	if (cur_seg != 0xffffffffu) {
		flush(cur_seg, false);
		if (FUSE)
			for (int i = tid; i < (NB << fbits); i += THREADS) {
				const uint32_t v = sh2[i];
				if (v) atomicAdd(&c.fused[i], v);
			}
	}
}

// The same pass for a grid that cannot keep many blocks per SM (the cooperative tail kernel,
// msb64_tail.cuh: three blocks of 256 threads, held down by the scatter pass' footprint).
// With so few warps the load -> count -> load rhythm of histogram_pass leaves HBM idle half
// of the time (measured: 2.6 instead of 1.3 ms per 2^30 keys), so here the keys of tile t + 1
// travel into shared memory by a bulk asynchronous copy (cp.async.bulk + mbarrier, issued by
// one thread, two stages) while the block counts tile t out of shared memory: every block
// has 32 KiB in flight all the time and nobody's registers wait for a load.
// Not fused, no level-0 origin: tail levels only.  Shared memory: HistStagedCfg::SMEM.
template <int BITS, int THREADS>
struct HistStagedCfg {
	static constexpr int NB = 1 << BITS;
	static constexpr size_t SMEM = 2 * size_t(TILE) * 8              // two stages of keys
				       + size_t(NB + 32 + 4) * 4         // histogram, dummy bins, OR / AND words
				       + 3 * 32                          // tile descriptors, three deep
				       + 16;                             // two mbarriers
};

struct HistDesc {
	uint32_t seg, begin, end, lo;
	uint32_t flags;   // HD_* bits
	uint32_t shift;
	uint32_t pad[2];
};
constexpr uint32_t HD_BUF = 1u, HD_READY = 2u, HD_WANT = 4u, HD_NONE = 8u;

template <int BITS, int THREADS>
__device__ __forceinline__ void histogram_pass_staged(const Ctx &c, const int level)
{
	using Cfg = HistStagedCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, ITEMS = TILE / THREADS;
	static_assert(ITEMS % 2 == 0, "tile is read as 16-byte pairs");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *stage = reinterpret_cast<uint64_t *>(smem_raw);            // [2][TILE]
	uint32_t *sh = reinterpret_cast<uint32_t *>(stage + 2 * TILE);       // [NB + 32]
	uint32_t *sbits = sh + NB + 32;                                      // [4]
	HistDesc *sdesc = reinterpret_cast<HistDesc *>(sbits + 4);           // [3]
	uint64_t *bar = reinterpret_cast<uint64_t *>(sdesc + 3);             // [2]

	const uint32_t tid = threadIdx.x;
	const uint32_t ntiles = c.ctl->ntiles[level];
	if (c.ctl->nready[level] == c.ctl->nsegs[level]) return;
	if (blockIdx.x >= ntiles) return;
	const Seg *segs = (level & 1) ? c.segs[1] : c.segs[0];
	const Tile *tiles = (level & 1) ? c.tiles[1] : c.tiles[0];
	uint32_t *hist = (level & 1) ? c.hist[1] : c.hist[0];
	SegBits *segbits = (level & 1) ? c.segbits[1] : c.segbits[0];
	const uint32_t G = gridDim.x;

	// thread 0 only
	auto fetch_desc = [&](uint32_t t, HistDesc *slot) {
		if (t >= ntiles) {
			slot->flags = HD_NONE;
			return;
		}
		const Tile tile = tiles[t];
		const Seg s = segs[tile.seg];
		slot->seg = tile.seg;
		slot->begin = s.begin;
		slot->end = s.begin + s.size;
		slot->lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		slot->flags = (s.buf ? HD_BUF : 0u) | ((s.flags & SEG_HIST_READY) ? HD_READY : 0u) |
			      ((s.flags & (SEG_WANT_BITS | SEG_HIST_READY)) == SEG_WANT_BITS ? HD_WANT : 0u);
		slot->shift = uint32_t(seg_shift(s.flags));
	};
	// slots of the tile's window a bulk copy fetches (0: it would leave the array -- plain loads)
	auto window = [&](uint32_t lo, uint32_t end) -> uint32_t {
		const uint32_t w = min(uint32_t(TILE), (end - lo + 1u) & ~1u);
		return lo + w <= c.end ? w : 0u;
	};
	auto start_copy = [&](const HistDesc &d, uint32_t st) {
		if (d.flags & (HD_READY | HD_NONE)) return;
		const uint32_t w = window(d.lo, d.end);
		if (!w) return;
		mbar_expect_tx(&bar[st], w * 8);
		bulk_copy_g2s(stage + st * TILE, ((d.flags & HD_BUF) ? c.keys[1] : c.keys[0]) + d.lo, w * 8, &bar[st]);
	};

	unsigned long long bor = 0, band = ~0ull;
	bool want = false;
	auto flush = [&](uint32_t seg) {
		if (want) {
			const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(bor));
			const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(bor >> 32));
			const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(band));
			const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(band >> 32));
			if (lane_id() == 0) {
				atomicOr(&sbits[0], olo);
				atomicOr(&sbits[1], ohi);
				atomicAnd(&sbits[2], alo);
				atomicAnd(&sbits[3], ahi);
			}
			bor = 0;
			band = ~0ull;
		}
		__syncthreads();
		for (int b = tid; b < NB; b += THREADS) {
			const uint32_t v = sh[b];
			if (v) {
				atomicAdd(&hist[size_t(seg) * NB + b], v);
				sh[b] = 0;
			}
		}
		if (want && tid == 0) {
			atomicOr(&segbits[seg].vor, ((unsigned long long) sbits[1] << 32) | sbits[0]);
			atomicAnd(&segbits[seg].vand, ((unsigned long long) sbits[3] << 32) | sbits[2]);
			sbits[0] = sbits[1] = 0u;
			sbits[2] = sbits[3] = 0xffffffffu;
		}
		__syncthreads();
	};

	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		fetch_desc(blockIdx.x, &sdesc[0]);
		fetch_desc(blockIdx.x + G, &sdesc[1]);
		start_copy(sdesc[0], 0);
	}
	for (int i = tid; i < NB + 32; i += THREADS) sh[i] = 0;
	if (tid < 4) sbits[tid] = tid < 2 ? 0u : 0xffffffffu;
	__syncthreads();

	uint32_t cur_seg = 0xffffffffu, parity = 0, slot = 0, st = 0;
	for (uint32_t t = blockIdx.x; t < ntiles; t += G, slot = slot == 2 ? 0 : slot + 1, st ^= 1u) {
		const HistDesc cur = sdesc[slot];
		// the other stage and the descriptor slot two ahead are free: everybody passed the
		// barrier that closes the previous tile
		if (tid == 0) {
			start_copy(sdesc[slot == 2 ? 0 : slot + 1], st ^ 1u);
			fetch_desc(t + 2 * G, &sdesc[slot == 0 ? 2 : slot - 1]);
		}
		if (cur.seg != cur_seg) {
			if (cur_seg != 0xffffffffu) flush(cur_seg);
			cur_seg = cur.seg;
		}
		want = cur.flags & HD_WANT;
		if (!(cur.flags & HD_READY)) {
			uint64_t *kin = stage + st * TILE;
			if (window(cur.lo, cur.end)) {
				mbar_wait(&bar[st], (parity >> st) & 1u);
				parity ^= 1u << st;
			} else {
				// the window would cross the end of the array (last tile only): plain loads
				const uint64_t *keys = (cur.flags & HD_BUF) ? c.keys[1] : c.keys[0];
				for (uint32_t i = tid; i < TILE; i += THREADS) {
					const uint32_t e = cur.lo + i;
					if (e >= cur.begin && e < cur.end) kin[i] = ld_stream_u64(keys + e);
				}
				__syncthreads();
			}
			const bool full = cur.lo >= cur.begin && cur.lo + TILE <= cur.end;
			uint64_t k[ITEMS];
			uint32_t validmask = (1u << ITEMS) - 1;
			const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(kin);
#pragma unroll
			for (int j = 0; j < ITEMS / 2; ++j) {
				const ulonglong2 v = k2[j * THREADS + tid];
				k[2 * j] = v.x;
				k[2 * j + 1] = v.y;
			}
			if (!full) {
				validmask = 0;
#pragma unroll
				for (int j = 0; j < ITEMS; ++j) {
					const uint32_t e = cur.lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
					validmask |= uint32_t(e >= cur.begin && e < cur.end) << j;
				}
			}
			if (want) {
#pragma unroll
				for (int j = 0; j < ITEMS; ++j)
					if ((validmask >> j) & 1u) {
						bor |= k[j];
						band &= k[j];
					}
			}
			hist_add_tile<ITEMS, NB, false>(sh, k, int(cur.shift), 0u, validmask);
		}
		__syncthreads();
	}
	if (cur_seg != 0xffffffffu) flush(cur_seg);
	if (tid == 0) {
		mbar_inval(&bar[0]);
		mbar_inval(&bar[1]);
	}
}

template <int BITS, int THREADS, bool FUSE>
__global__ void __launch_bounds__(THREADS, 4)
histogram_kernel(const Ctx c, const int level, const uint32_t origin, const int fbits)
{
	histogram_pass<BITS, THREADS, FUSE>(c, level, origin, fbits);
}

} // namespace msb64
