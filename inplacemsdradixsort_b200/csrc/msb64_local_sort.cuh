// msb64_local_sort.cuh -- finish of small buckets entirely in shared memory (the role of
// insertsort / combsort and the in-cache partition_ip levels, msb_64.c:126-149, 980-1005,
// 740-770).  Algorithmic traffic: 32 bytes per pair (16 read + 16 written), once.
//
// A unit is a run of neighbouring buckets of at most LOCAL_CAP pairs.  Its keys agree
// on some prefix and may differ anywhere below; the block
//   1. loads keys and rids (coalesced) and reduces OR and AND of the keys: the bits in
//      which the unit's keys differ are exactly OR & ~AND;
//   2. counting-sorts on the top `b` differing bits with 1-2 bins per key: a shared
//      atomic gives every key its arrival rank inside its bin, a block scan over the
//      bin counts (16-byte shared loads, 4 bins each) gives the bin bases, and every
//      pair goes straight to base + arrival rank in the staging arrays;
//   3. bins that hold several different keys are rare and short (Poisson with mean
//      <= 1): the first arrival files the bin in a dense list, and afterwards ONE thread
//      per listed bin orders its few pairs in place by insertion -- all lanes of a warp
//      work on colliding bins, instead of every key looping over its bin mates behind
//      a divergent branch.  Bins longer than LOCAL_SERIAL_MAX (adversarial bit
//      patterns; random keys never produce them) are finished by a block-wide bitonic
//      network;
//   4. writes keys and rids from the staging arrays to the caller's arrays, coalesced.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

#ifndef MSB64_LOCAL_ITEMS
#define MSB64_LOCAL_ITEMS 8
#endif
constexpr int LOCAL_ITEMS = MSB64_LOCAL_ITEMS;   // pairs per thread
constexpr int LOCAL_THREADS = LOCAL_CAP / LOCAL_ITEMS;
#ifndef MSB64_LOCAL_MINB
#define MSB64_LOCAL_MINB 2
#endif
#ifndef MSB64_LOCAL_GENERAL_MINB
#define MSB64_LOCAL_GENERAL_MINB MSB64_LOCAL_MINB
#endif
constexpr int LOCAL_MINB = MSB64_LOCAL_MINB;     // resident blocks per SM the register budget is cut for
constexpr int LOCAL_OWNERS = LOCAL_THREADS >= 512 ? 512 : 256;   // threads that own bins in the scan
constexpr int LOCAL_LOG_OWNERS = LOCAL_OWNERS == 512 ? 9 : 8;
static_assert(LOCAL_THREADS % 32 == 0 && LOCAL_THREADS >= LOCAL_OWNERS && LOCAL_THREADS <= 1024, "block shape");
#ifndef MSB64_LOCAL_BITS
#define MSB64_LOCAL_BITS 12
#endif
constexpr int LOCAL_BITS = MSB64_LOCAL_BITS;    // at most 2^LOCAL_BITS bins
#ifndef MSB64_LOCAL_EXTRA
#define MSB64_LOCAL_EXTRA 0
#endif
constexpr int LOCAL_EXTRA_BITS = MSB64_LOCAL_EXTRA;   // bins per key: 2^EXTRA .. 2^(EXTRA+1)
constexpr int LOCAL_LPER = LOCAL_BITS - LOCAL_LOG_OWNERS;      // log2(bins per thread): 8 or 16 bins, as chunks of 4
constexpr int LOCAL_CHUNKS = (1 << LOCAL_LPER) / 4;
static_assert(LOCAL_LPER >= 2 && LOCAL_LPER <= 5, "bin layout");
constexpr uint32_t LOCAL_NBINS = 1u << LOCAL_BITS;
constexpr uint32_t LOCAL_SERIAL_MAX = 16;       // bins up to this size: ordered by one thread
constexpr uint32_t LOCAL_LIST_MAX = LOCAL_CAP / 2;                       // bins with >= 2 keys
constexpr uint32_t LOCAL_BIG_MAX = LOCAL_CAP / (LOCAL_SERIAL_MAX + 1) + 2;
constexpr size_t LOCAL_SMEM = size_t(LOCAL_CAP) * 16 + (size_t(LOCAL_NBINS) + 32) * 4
			      + LOCAL_LIST_MAX * 4 + LOCAL_BIG_MAX * 4 + 64 * 4
			      + 2 * (LOCAL_THREADS / 32) * 8;

// OR and AND of one 64-bit value per thread over the block, in two steps around a barrier the
// caller places (both local-sort kernels do other work between them): every warp leaves its
// partial results in wred[0 .. WARPS) (OR) and wred[WARPS .. 2 WARPS) (AND) ...
template <int WARPS>
__device__ __forceinline__ void or_and_warp_to_shared(uint64_t vor, uint64_t vand, uint64_t *wred)
{
	const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(vor));
	const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(vor >> 32));
	const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(vand));
	const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(vand >> 32));
	if (lane_id() == 0) {
		wred[threadIdx.x >> 5] = (uint64_t(ohi) << 32) | olo;
		wred[WARPS + (threadIdx.x >> 5)] = (uint64_t(ahi) << 32) | alo;
	}
}
// ... and, behind the barrier, every warp folds them (the result is the same on all threads)
template <int WARPS>
__device__ __forceinline__ void or_and_from_shared(const uint64_t *wred, uint64_t *vor, uint64_t *vand)
{
	static_assert(WARPS <= 32, "one lane per warp result");
	const uint32_t lane = lane_id();
	const uint64_t o = lane < WARPS ? wred[lane] : 0ull;
	const uint64_t a = lane < WARPS ? wred[WARPS + lane] : ~0ull;
	const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(o));
	const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(o >> 32));
	const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(a));
	const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(a >> 32));
	*vor = (uint64_t(ohi) << 32) | olo;
	*vand = (uint64_t(ahi) << 32) | alo;
}

// Ascending compare-exchange network for any length (bitonic merges with the first
// step mirrored, so that the missing tail behaves like +infinity).
__device__ __forceinline__ void block_bitonic(uint64_t *k, uint64_t *r, const uint32_t n)
{
	for (uint32_t width = 2; (width >> 1) < n; width <<= 1) {
		for (uint32_t j = width >> 1; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
				const uint32_t p = (j == (width >> 1)) ? (i ^ (width - 1)) : (i ^ j);
				if (p > i && p < n) {
					const uint64_t a = k[i], b = k[p];
					if (a > b) {
						k[i] = b;
						k[p] = a;
						const uint64_t t = r[i];
						r[i] = r[p];
						r[p] = t;
					}
				}
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(LOCAL_THREADS, MSB64_LOCAL_GENERAL_MINB)
local_sort_kernel(const Ctx c, const uint64_t base_key)
{
	constexpr int THREADS = LOCAL_THREADS, ITEMS = LOCAL_ITEMS, WARPS = THREADS / 32;
	constexpr int OWNERS = LOCAL_OWNERS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);        // [LOCAL_CAP]
	uint64_t *srids = skeys + LOCAL_CAP;                             // [LOCAL_CAP]
	uint32_t *bins = reinterpret_cast<uint32_t *>(srids + LOCAL_CAP);// [LOCAL_NBINS + 32]
	uint32_t *list = bins + LOCAL_NBINS + 32;                        // short bins to order: base | size << 16
	uint32_t *big = list + LOCAL_LIST_MAX;                           // long bins:           base | size << 16
	uint32_t *scratch = big + LOCAL_BIG_MAX;                         // [64]
	uint64_t *wred = reinterpret_cast<uint64_t *>(scratch + 64);     // [2 * WARPS] OR, AND per warp
	__shared__ uint32_t s_nbig;

	const uint32_t tid = threadIdx.x, lane = lane_id();
	const uint32_t nunits = min(c.ctl->nslow, c.max_units);          // this kernel's units sit at the back
	const Unit *units = c.units + c.max_units - 1;                   // unit i is units[-i]

	// The loop is software-pipelined: the pairs of the NEXT unit are requested from HBM as
	// soon as the current unit's pairs have left the registers for shared memory (after step
	// 3a), so that their latency hides behind steps 3b and 4.
	uint64_t k[ITEMS], r[ITEMS];
	// Rows (THREADS consecutive slots) past the unit's end are skipped with block-uniform
	// branches; the slots of the last row past the end re-read the last pair (no
	// divergence, and harmless for the OR / AND reductions).
	auto load_unit = [&](const Unit &x) {
		const uint64_t *src_keys = (x.buf ? c.keys[1] : c.keys[0]) + x.begin;
		const uint64_t *src_rids = (x.buf ? c.rids[1] : c.rids[0]) + x.begin;
		const uint32_t nrows = (x.size + THREADS - 1) / THREADS;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < nrows) k[j] = ld_stream_u64(src_keys + min(uint32_t(j * THREADS) + tid, x.size - 1));
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < nrows) r[j] = ld_stream_u64(src_rids + min(uint32_t(j * THREADS) + tid, x.size - 1));
	};
	Unit next = Unit{0u, 1u, 0u, 0u};
	if (blockIdx.x < nunits) {
		next = *(units - blockIdx.x);
		load_unit(next);
	}

	for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
		// 1. the unit's pairs are in (or on their way to) the registers
		const Unit un = next;
		const bool more = u + gridDim.x < nunits;
		const uint32_t begin = un.begin;
		const uint32_t size = un.size;
		const uint32_t rows = (size + THREADS - 1) / THREADS;
		// the bin table is free here (the previous unit is done with it)
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			for (uint32_t i = tid; i < (LOCAL_NBINS + 32) / 4; i += THREADS)
				b4[i] = make_uint4(0u, 0u, 0u, 0u);
		}
		if (tid == 0) s_nbig = 0;
		// binning value: key minus the unit's origin (monotone; see unit_origin)
		// units filed at level 0 of a range-limited sort carry digits relative to base_key
		const uint64_t origin = unit_origin_key(un.origin) + ((un.origin & UNIT_LEVEL0) ? base_key : 0ull);
		uint64_t vor = k[0] - origin, vand = vor;
#pragma unroll
		for (int j = 1; j < ITEMS; ++j)
			if (j < rows) {
				vor |= k[j] - origin;
				vand &= k[j] - origin;
			}
		or_and_warp_to_shared<WARPS>(vor, vand, wred);
		__syncthreads();
		or_and_from_shared<WARPS>(wred, &vor, &vand);
		const uint64_t diff = vor & ~vand;

		if (diff == 0) {
			// all keys equal: nothing to order, only bring the pairs home
			if (un.buf != 0) {
#pragma unroll
				for (int j = 0; j < ITEMS; ++j) {
					const uint32_t i = j * THREADS + tid;
					if (i < size) {
						st_stream_u64(c.keys[0] + begin + i, k[j]);
						st_stream_u64(c.rids[0] + begin + i, r[j]);
					}
				}
			}
			__syncthreads();
			if (more) {
				next = *(units - (u + gridDim.x));
				load_unit(next);
			}
			continue;
		}

		// 2. counting sort on the top differing bits, 1-2 bins per key
		const int top = 63 - __clzll(diff);                      // highest differing bit
		int b = 32 - __clz(size - 1);                            // ceil(log2(size)), size >= 2 here
		b = min(max(b + LOCAL_EXTRA_BITS, 5), LOCAL_BITS);
		b = min(b, top + 1);
		const int shift = top + 1 - b;
		const uint32_t nb = 1u << b, dmask = nb - 1;
		// thread t owns the PER (8 or 16) consecutive digits t*PER .. t*PER+PER-1 and keeps
		// them as 16-byte chunks of four at chunk indices c*OWNERS + t: the scan reads and
		// writes them with conflict-free 16-byte accesses
		constexpr int LPER = LOCAL_LPER, CH = LOCAL_CHUNKS;
#ifdef MSB64_NATURAL_BINS
#define MSB64_BIN_SLOT(d) (d)
#define MSB64_BIN_CHUNK(ch, t) ((t) * CH + (ch))
#else
#define MSB64_BIN_SLOT(d) ((((((d) >> 2) & (CH - 1)) * OWNERS + ((d) >> LPER)) << 2) | ((d) & 3u))
#define MSB64_BIN_CHUNK(ch, t) ((ch) * OWNERS + (t))
#endif
		// do the digit bits cover every differing bit?  then equal digit = equal key
		const bool resolved = (diff & ((1ull << shift) - 1)) == 0;

		// branch-free inside a row (see tile_ranks in msb64_scatter.cuh): slots past the
		// unit's end count into per-lane dummy bins behind the table
		uint32_t rk[ITEMS / 2];                 // arrival ranks, two 16-bit values per register
#pragma unroll
		for (int j = 0; j < ITEMS / 2; ++j) rk[j] = 0;
		uint32_t *dummy = bins + LOCAL_NBINS + lane;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) {
				const uint32_t i = j * THREADS + tid;
				const uint32_t d = uint32_t((k[j] - origin) >> shift) & dmask;
				uint32_t *slot = i < size ? &bins[MSB64_BIN_SLOT(d)] : dummy;
				rk[j >> 1] |= atomicAdd(slot, 1u) << (16 * (j & 1));
			}
		__syncthreads();
		// exclusive scan over bins in digit order; every bin becomes base | count << 16.
		// The same scan numbers the bins that hold 2..LOCAL_SERIAL_MAX keys (count of those
		// in the upper half of the scanned word): they are filed in `list` without atomics.
		uint32_t nlist;
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			const bool own = tid < OWNERS && (tid << LPER) < nb;
			uint32_t cn[4 * CH];
#pragma unroll
			for (int ch = 0; ch < CH; ++ch) {
				uint4 v = make_uint4(0u, 0u, 0u, 0u);
				if (own) v = b4[MSB64_BIN_CHUNK(ch, tid)];
				cn[4 * ch] = v.x;
				cn[4 * ch + 1] = v.y;
				cn[4 * ch + 2] = v.z;
				cn[4 * ch + 3] = v.w;
			}
			uint32_t sum = 0;
#pragma unroll
			for (int q = 0; q < 4 * CH; ++q) {
				sum += cn[q];
				if (!resolved && cn[q] - 2u <= LOCAL_SERIAL_MAX - 2u) sum += 1u << 16;
			}
			uint32_t total;
			const uint32_t ex = block_exclusive_scan<THREADS>(sum, scratch, &total);
			nlist = total >> 16;
			if (own) {
				uint32_t base = ex & 0xffffu, at = ex >> 16;
#pragma unroll
				for (int q = 0; q < 4 * CH; ++q) {
					const uint32_t o = base | (cn[q] << 16);
					base += cn[q];
					if (!resolved && cn[q] >= 2u) {
						if (cn[q] <= LOCAL_SERIAL_MAX) list[at++] = o;
						else big[atomicAdd(&s_nbig, 1u)] = uint32_t((MSB64_BIN_CHUNK(q >> 2, tid) << 2) | (q & 3));   // where the bin sits in the table
					}
					cn[q] = o;
				}
#pragma unroll
				for (int ch = 0; ch < CH; ++ch)
					b4[MSB64_BIN_CHUNK(ch, tid)] = make_uint4(cn[4 * ch], cn[4 * ch + 1], cn[4 * ch + 2], cn[4 * ch + 3]);
			}
		}
		__syncthreads();

		// 3a. every pair to bin base + arrival rank
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) {
				const uint32_t i = j * THREADS + tid;
				if (i < size) {
					const uint32_t d = uint32_t((k[j] - origin) >> shift) & dmask;
					const uint32_t p = (bins[MSB64_BIN_SLOT(d)] & 0xffffu) + ((rk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
					skeys[p] = k[j];
					srids[p] = r[j];
				}
			}
		__syncthreads();
		const uint32_t nbig = s_nbig;     // read before the next barrier: thread 0 resets it for the next unit after it
		if (nbig) {
			// long bins (duplicate-heavy or adversarial keys): a bin whose keys are all equal
			// needs no ordering.  Every key of a long bin compares itself with the bin's first
			// key and flags the bin (bit 31 of its table entry) when it differs.
#pragma unroll
			for (int j = 0; j < ITEMS; ++j)
				if (j < rows) {
					const uint32_t i = j * THREADS + tid;
					if (i < size) {
						const uint32_t d = uint32_t((k[j] - origin) >> shift) & dmask;
						const uint32_t pk = bins[MSB64_BIN_SLOT(d)];
						if (((pk >> 16) & 0x7fffu) > LOCAL_SERIAL_MAX && !(pk >> 31) &&
						    skeys[pk & 0xffffu] != k[j])
							atomicOr(&bins[MSB64_BIN_SLOT(d)], 1u << 31);
					}
				}
			__syncthreads();
			if (tid < nbig) big[tid] = bins[big[tid]];      // table position -> base | size << 16 | differs << 31
		}
		// the registers are free: request the next unit's pairs now
		if (more) {
			next = *(units - (u + gridDim.x));
			load_unit(next);
		}
		// 3b. one thread per short colliding bin.  Up to four pairs (nearly all of them):
		//     loaded at once, ordered in registers by a 5-comparator network, stored back --
		//     one shared-memory round trip instead of a chain of dependent ones.  Missing
		//     slots hold the largest key and are never moved down (exchanges are strict).
		for (uint32_t q = tid; q < nlist; q += THREADS) {
			const uint32_t pk = list[q];
			uint64_t *bk = skeys + (pk & 0xffffu), *br = srids + (pk & 0xffffu);
			const uint32_t cnt = pk >> 16;
			if (cnt == 2) {
				// the common case: rids are touched only when the two keys are out of order
				const uint64_t a0 = bk[0], a1 = bk[1];
				if (a0 > a1) {
					const uint64_t b0 = br[0], b1 = br[1];
					bk[0] = a1;
					bk[1] = a0;
					br[0] = b1;
					br[1] = b0;
				}
				continue;
			}
			if (cnt <= 4) {
				uint64_t a0 = bk[0], a1 = bk[1], b0 = br[0], b1 = br[1];
				uint64_t a2 = ~0ull, a3 = ~0ull, b2 = 0, b3 = 0;
				if (cnt > 2) {
					a2 = bk[2];
					b2 = br[2];
				}
				if (cnt > 3) {
					a3 = bk[3];
					b3 = br[3];
				}
#define MSB64_CE(x, y, rx, ry)                                                   \
	{                                                                         \
		const bool sw = x > y;                                            \
		const uint64_t tx = sw ? y : x, ty = sw ? x : y;                  \
		const uint64_t trx = sw ? ry : rx, try_ = sw ? rx : ry;           \
		x = tx; y = ty; rx = trx; ry = try_;                              \
	}
				MSB64_CE(a0, a1, b0, b1)
				MSB64_CE(a2, a3, b2, b3)
				MSB64_CE(a0, a2, b0, b2)
				MSB64_CE(a1, a3, b1, b3)
				MSB64_CE(a1, a2, b1, b2)
#undef MSB64_CE
				bk[0] = a0; br[0] = b0;
				bk[1] = a1; br[1] = b1;
				if (cnt > 2) {
					bk[2] = a2;
					br[2] = b2;
				}
				if (cnt > 3) {
					bk[3] = a3;
					br[3] = b3;
				}
				continue;
			}
			for (uint32_t i = 1; i < cnt; ++i) {
				const uint64_t key = bk[i];
				uint32_t at = i;
				uint64_t prev = bk[at - 1];
				if (prev <= key) continue;
				const uint64_t rid = br[i];
				do {
					bk[at] = prev;
					br[at] = br[at - 1];
					--at;
					if (at == 0) break;
					prev = bk[at - 1];
				} while (prev > key);
				bk[at] = key;
				br[at] = rid;
			}
		}
		__syncthreads();
		// 3c. long bins whose keys are not all equal (adversarial bit patterns): block-wide network
		for (uint32_t q = 0; q < nbig; ++q) {
			const uint32_t pk = big[q];
			if (pk >> 31) block_bitonic(skeys + (pk & 0xffffu), srids + (pk & 0xffffu), (pk >> 16) & 0x7fffu);
		}

		// 4. home (the next unit's first barrier orders these reads before its stores)
		for (uint32_t i = tid; i < size; i += THREADS) {
			st_stream_u64(c.keys[0] + begin + i, skeys[i]);
			st_stream_u64(c.rids[0] + begin + i, srids[i]);
		}
#undef MSB64_BIN_SLOT
#undef MSB64_BIN_CHUNK
	}
}

} // namespace msb64
