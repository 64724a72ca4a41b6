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

constexpr int LOCAL_THREADS = 512;
constexpr int LOCAL_ITEMS = LOCAL_CAP / LOCAL_THREADS;
constexpr int LOCAL_BITS = 12;                  // at most 4096 bins
constexpr uint32_t LOCAL_NBINS = 1u << LOCAL_BITS;
constexpr uint32_t LOCAL_SERIAL_MAX = 16;       // bins up to this size: ordered by one thread
constexpr uint32_t LOCAL_LIST_MAX = LOCAL_CAP / 2;                       // bins with >= 2 keys
constexpr uint32_t LOCAL_BIG_MAX = LOCAL_CAP / (LOCAL_SERIAL_MAX + 1) + 2;
constexpr size_t LOCAL_SMEM = size_t(LOCAL_CAP) * 16 + (size_t(LOCAL_NBINS) + 32) * 4
			      + LOCAL_LIST_MAX * 4 + LOCAL_BIG_MAX * 4 + 64 * 4
			      + 2 * (LOCAL_THREADS / 32) * 8;

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

__global__ void __launch_bounds__(LOCAL_THREADS, 2)
local_sort_kernel(const Ctx c)
{
	constexpr int THREADS = LOCAL_THREADS, ITEMS = LOCAL_ITEMS, WARPS = THREADS / 32;
	static_assert(THREADS == 512, "the bin layout assumes 512 threads");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);        // [LOCAL_CAP]
	uint64_t *srids = skeys + LOCAL_CAP;                             // [LOCAL_CAP]
	uint32_t *bins = reinterpret_cast<uint32_t *>(srids + LOCAL_CAP);// [LOCAL_NBINS + 32]
	uint32_t *list = bins + LOCAL_NBINS + 32;                        // short bins to order: base | size << 16
	uint32_t *big = list + LOCAL_LIST_MAX;                           // long bins:           base | size << 16
	uint32_t *scratch = big + LOCAL_BIG_MAX;                         // [64]
	uint64_t *wred = reinterpret_cast<uint64_t *>(scratch + 64);     // [2 * WARPS] OR, AND per warp
	__shared__ uint32_t s_nbig;

	const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
	const uint32_t nunits = min(c.ctl->nunits, c.max_units);

	for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
		const Unit un = c.units[u];
		const uint64_t *src_keys = (un.buf ? c.keys[1] : c.keys[0]) + un.begin;
		const uint64_t *src_rids = (un.buf ? c.rids[1] : c.rids[0]) + un.begin;
		uint64_t *dst_keys = c.keys[0] + un.begin, *dst_rids = c.rids[0] + un.begin;
		const uint32_t size = un.size;

		// 1. load.  Rows (THREADS consecutive slots) past the unit's end are skipped with
		//    block-uniform branches; the slots of the last row past the end re-read the
		//    last pair (no divergence, and harmless for the OR / AND reductions).
		const uint32_t rows = (size + THREADS - 1) / THREADS;
		uint64_t k[ITEMS], r[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) k[j] = ld_stream_u64(src_keys + min(uint32_t(j * THREADS) + tid, size - 1));
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) r[j] = ld_stream_u64(src_rids + min(uint32_t(j * THREADS) + tid, size - 1));
		// the bin table is free here (the previous unit is done with it)
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			for (uint32_t i = tid; i < (LOCAL_NBINS + 32) / 4; i += THREADS)
				b4[i] = make_uint4(0u, 0u, 0u, 0u);
		}
		if (tid == 0) s_nbig = 0;
		// binning value: key minus the unit's origin (monotone; see unit_origin)
		const uint64_t origin = unit_origin_key(un.origin);
		uint64_t vor = k[0] - origin, vand = vor;
#pragma unroll
		for (int j = 1; j < ITEMS; ++j)
			if (j < rows) {
				vor |= k[j] - origin;
				vand &= k[j] - origin;
			}
		{
			const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(vor));
			const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(vor >> 32));
			const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(vand));
			const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(vand >> 32));
			if (lane == 0) {
				wred[warp] = (uint64_t(ohi) << 32) | olo;
				wred[WARPS + warp] = (uint64_t(ahi) << 32) | alo;
			}
		}
		__syncthreads();
		uint64_t diff;
		{
			static_assert(WARPS <= 32, "one lane per warp result");
			const uint64_t o = lane < WARPS ? wred[lane] : 0ull;
			const uint64_t a = lane < WARPS ? wred[WARPS + lane] : ~0ull;
			const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(o));
			const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(o >> 32));
			const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(a));
			const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(a >> 32));
			diff = ((uint64_t(ohi) << 32) | olo) & ~((uint64_t(ahi) << 32) | alo);
		}

		if (diff == 0) {
			// all keys equal: nothing to order, only bring the pairs home
			if (un.buf != 0) {
#pragma unroll
				for (int j = 0; j < ITEMS; ++j) {
					const uint32_t i = j * THREADS + tid;
					if (i < size) {
						st_stream_u64(dst_keys + i, k[j]);
						st_stream_u64(dst_rids + i, r[j]);
					}
				}
			}
			__syncthreads();
			continue;
		}

		// 2. counting sort on the top differing bits, 1-2 bins per key
		const int top = 63 - __clzll(diff);                      // highest differing bit
		int b = 32 - __clz(size - 1);                            // ceil(log2(size)), size >= 2 here
		b = min(max(b, 5), LOCAL_BITS);
		b = min(b, top + 1);
		const int shift = top + 1 - b;
		const uint32_t nb = 1u << b, dmask = nb - 1;
		// thread t owns the 8 consecutive digits 8t .. 8t+7 and keeps them as two 16-byte
		// chunks at chunk indices t and THREADS + t: the scan reads and writes them with
		// conflict-free 16-byte accesses
#define MSB64_BIN_SLOT(d) ((((d) & 4u) << 9) | (((d) >> 3) << 2) | ((d) & 3u))
		static_assert(LOCAL_NBINS == 8 * THREADS, "two chunks of four bins per thread");
		// do the digit bits cover every differing bit?  then equal digit = equal key
		const bool resolved = (diff & ((1ull << shift) - 1)) == 0;

		// branch-free inside a row (see tile_ranks in msb64_scatter.cuh): slots past the
		// unit's end count into per-lane dummy bins behind the table
		uint32_t rank[ITEMS];
		uint32_t *dummy = bins + LOCAL_NBINS + lane;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) {
				const uint32_t i = j * THREADS + tid;
				const uint32_t d = uint32_t((k[j] - origin) >> shift) & dmask;
				uint32_t *slot = i < size ? &bins[MSB64_BIN_SLOT(d)] : dummy;
				rank[j] = atomicAdd(slot, 1u);
			}
		__syncthreads();
		// exclusive scan over bins in digit order; every bin becomes base | count << 16.
		// The same scan numbers the bins that hold 2..LOCAL_SERIAL_MAX keys (count of those
		// in the upper half of the scanned word): they are filed in `list` without atomics.
		uint32_t nlist;
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			const bool own = (tid << 3) < nb;
			uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
			if (own) {
				v0 = b4[tid];
				v1 = b4[THREADS + tid];
			}
			const uint32_t cn[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
			uint32_t sum = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				sum += cn[q];
				if (!resolved && cn[q] - 2u <= LOCAL_SERIAL_MAX - 2u) sum += 1u << 16;
			}
			uint32_t total;
			const uint32_t ex = block_exclusive_scan<THREADS>(sum, scratch, &total);
			nlist = total >> 16;
			if (own) {
				uint32_t base = ex & 0xffffu, at = ex >> 16;
				uint32_t o[8];
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					o[q] = base | (cn[q] << 16);
					base += cn[q];
					if (!resolved && cn[q] >= 2u) {
						if (cn[q] <= LOCAL_SERIAL_MAX) list[at++] = o[q];
						else big[atomicAdd(&s_nbig, 1u)] = o[q];
					}
				}
				b4[tid] = make_uint4(o[0], o[1], o[2], o[3]);
				b4[THREADS + tid] = make_uint4(o[4], o[5], o[6], o[7]);
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
					const uint32_t p = (bins[MSB64_BIN_SLOT(d)] & 0xffffu) + rank[j];
					skeys[p] = k[j];
					srids[p] = r[j];
				}
			}
		__syncthreads();
		// 3b. one thread per short colliding bin: insertion sort in place
		const uint32_t nbig = s_nbig;     // read before the next barrier: thread 0 resets it for the next unit after it
		for (uint32_t q = tid; q < nlist; q += THREADS) {
			const uint32_t pk = list[q];
			uint64_t *bk = skeys + (pk & 0xffffu), *br = srids + (pk & 0xffffu);
			const uint32_t cnt = pk >> 16;
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
		for (uint32_t q = 0; q < nbig; ++q) {
			const uint32_t pk = big[q];
			block_bitonic(skeys + (pk & 0xffffu), srids + (pk & 0xffffu), pk >> 16);
		}

		// 4. home (the next unit's first barrier orders these reads before its stores)
		for (uint32_t i = tid; i < size; i += THREADS) {
			st_stream_u64(dst_keys + i, skeys[i]);
			st_stream_u64(dst_rids + i, srids[i]);
		}
#undef MSB64_BIN_SLOT
	}
}

} // namespace msb64
