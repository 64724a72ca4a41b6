// msb64_local_sort.cuh -- finish of small buckets entirely in shared memory (the role of
// insertsort / combsort and the in-cache partition_ip levels, msb_64.c:126-149, 980-1005,
// 740-770).  Algorithmic traffic: 32 bytes per pair (16 read + 16 written), once.
//
// A unit is a run of neighbouring buckets of at most LOCAL_CAP pairs.  Its keys agree
// on some prefix and may differ anywhere below; the block
//   1. loads keys and rids (coalesced) and ORs together key ^ first_key: the set bits
//      are exactly the bit positions in which the unit's keys differ;
//   2. counting-sorts on the top `b` differing bits with 2-4 bins per key (shared
//      atomics give every key its arrival rank inside its bin, a block scan gives the
//      bin bases);
//   3. resolves bins that hold several different keys without moving data: the keys of
//      such bins are parked in bin order, then every key counts the keys of its own bin
//      that precede it (all lanes busy, no serial insertion loops); bins longer than
//      LOCAL_RANK_MAX -- adversarial bit patterns, random keys never produce them -- are
//      finished by a block-wide bitonic network;
//   4. writes keys and rids to their final slots in shared memory and from there to
//      the caller's arrays, coalesced.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

constexpr int LOCAL_THREADS = 256;
constexpr int LOCAL_ITEMS = LOCAL_CAP / LOCAL_THREADS;
constexpr int LOCAL_BITS = 13;                  // at most 8192 bins
constexpr uint32_t LOCAL_RANK_MAX = 32;         // bins up to this size: rank by counting
constexpr uint32_t LOCAL_BIG_MAX = LOCAL_CAP / LOCAL_RANK_MAX + 2;
constexpr size_t LOCAL_SMEM = size_t(LOCAL_CAP) * 16 + ((size_t(1) << LOCAL_BITS) + 32) * 4
			      + LOCAL_BIG_MAX * 4 + 64 * 4 + 16 * 8;

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
	constexpr int THREADS = LOCAL_THREADS, ITEMS = LOCAL_ITEMS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);        // [LOCAL_CAP]
	uint64_t *srids = skeys + LOCAL_CAP;                             // [LOCAL_CAP]
	uint32_t *bins = reinterpret_cast<uint32_t *>(srids + LOCAL_CAP);// [1 << LOCAL_BITS]
	uint32_t *big = bins + (1u << LOCAL_BITS) + 32;                  // long bins: base | size << 16
	uint32_t *scratch = big + LOCAL_BIG_MAX;                         // [64]
	uint64_t *wor = reinterpret_cast<uint64_t *>(scratch + 64);      // [16]
	__shared__ uint32_t s_nbig;

	const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
	const uint32_t nunits = min(c.ctl->nunits, c.max_units);

	for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
		const Unit un = c.units[u];
		const uint64_t *src_keys = (un.buf ? c.keys[1] : c.keys[0]) + un.begin;
		const uint64_t *src_rids = (un.buf ? c.rids[1] : c.rids[0]) + un.begin;
		uint64_t *dst_keys = c.keys[0] + un.begin, *dst_rids = c.rids[0] + un.begin;
		const uint32_t size = un.size;

		// 1. load; which bits differ?
		uint64_t k[ITEMS], r[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			k[j] = i < size ? ld_stream_u64(src_keys + i) : 0;
		}
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			r[j] = i < size ? ld_stream_u64(src_rids + i) : 0;
		}
		if (tid == 0) {
			wor[8] = k[0];
			s_nbig = 0;
		}
		__syncthreads();
		const uint64_t k0 = wor[8];
		uint64_t diff = 0;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (uint32_t(j * THREADS + tid) < size) diff |= k[j] ^ k0;
		const uint32_t dlo = __reduce_or_sync(0xffffffffu, uint32_t(diff));
		const uint32_t dhi = __reduce_or_sync(0xffffffffu, uint32_t(diff >> 32));
		if (lane == 0) wor[warp] = (uint64_t(dhi) << 32) | dlo;
		__syncthreads();
		diff = 0;
#pragma unroll
		for (int w = 0; w < THREADS / 32; ++w) diff |= wor[w];

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

		// 2. counting sort on the top differing bits, 2-4 bins per key
		const int top = 63 - __clzll(diff);                      // highest differing bit
		int b = 32 - __clz(size - 1);                            // ceil(log2(size))
		b = min(max(b + 1, 5), LOCAL_BITS);
		b = min(b, top + 1);
		const int shift = top + 1 - b;
		const uint32_t nb = 1u << b, dmask = nb - 1;
		// bin table transposed so that the scan below is bank-conflict free: thread t owns
		// the `per` consecutive digits t*per .. t*per+per-1 and keeps them at q*THREADS + t
		const int lper = max(b - 8, 0);                          // log2(bins per thread)
		const uint32_t per = 1u << lper, pmask = per - 1;
#define MSB64_BIN_SLOT(d) ((((d) & pmask) << 8) | ((d) >> lper))
		// do the digit bits cover every differing bit?  then equal digit = equal key
		const bool resolved = (diff & ((1ull << shift) - 1)) == 0;

		static_assert(THREADS == 256, "MSB64_BIN_SLOT assumes 256 threads");
		for (uint32_t i = tid; i < max(nb, uint32_t(THREADS)); i += THREADS) bins[i] = 0;
		__syncthreads();
		// branch-free (see tile_ranks in msb64_scatter.cuh): slots past the unit's end count
		// into per-lane dummy bins behind the table
		uint32_t rank[ITEMS];
		uint32_t *dummy = bins + (1u << LOCAL_BITS) + lane;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			const uint32_t d = uint32_t(k[j] >> shift) & dmask;
			uint32_t *slot = i < size ? &bins[MSB64_BIN_SLOT(d)] : dummy;
			rank[j] = atomicAdd(slot, 1u);
		}
		__syncthreads();
		// exclusive scan over bins in digit order; pack base | count << 16
		{
			uint32_t sum = 0;
			for (uint32_t q = 0; q < per; ++q) sum += bins[q * THREADS + tid];
			uint32_t total;
			uint32_t base = block_exclusive_scan<THREADS>(sum, scratch, &total);
			for (uint32_t q = 0; q < per; ++q) {
				const uint32_t cnt = bins[q * THREADS + tid];
				bins[q * THREADS + tid] = base | (cnt << 16);
				base += cnt;
			}
		}
		__syncthreads();

		// 3a. park the keys of bins that need ordering (bin base + arrival rank);
		//     rank[] becomes base | count << 13 | min(arrival rank, 63) << 26
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < size) {
				const uint32_t d = uint32_t(k[j] >> shift) & dmask;
				const uint32_t pk = bins[MSB64_BIN_SLOT(d)];
				const uint32_t base = pk & 0xffffu, cnt = pk >> 16;
				if (!resolved && cnt > 1) skeys[base + rank[j]] = k[j];
				rank[j] = (base + (resolved || cnt == 1 || cnt > LOCAL_RANK_MAX ? rank[j] : 0u))
					  | (cnt << 13) | (min(rank[j], 63u) << 26);
			}
		}
		__syncthreads();
		// 3b. final slot = bin base + number of keys of the bin that go before this one
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < size) {
				const uint32_t base = rank[j] & 0x1fffu, cnt = (rank[j] >> 13) & 0x1fffu;
				const uint32_t arrival = rank[j] >> 26;
				uint32_t slot = base;                      // already final unless ...
				if (!resolved && cnt > 1) {
					if (cnt <= LOCAL_RANK_MAX) {
						uint32_t before = 0;
						for (uint32_t q = 0; q < cnt; ++q) {
							const uint64_t other = skeys[base + q];
							before += (other < k[j]) || (other == k[j] && q < arrival);
						}
						slot = base + before;
					} else if (arrival == 0) {
						big[atomicAdd(&s_nbig, 1u)] = (base) | (cnt << 16);
					}
				}
				rank[j] = slot;
			}
		}
		__syncthreads();
		// 4a. final slots
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < size) {
				skeys[rank[j]] = k[j];
				srids[rank[j]] = r[j];
			}
		}
		__syncthreads();
		const uint32_t nbig = s_nbig;
		for (uint32_t q = 0; q < nbig; ++q) {
			const uint32_t pk = big[q];
			block_bitonic(skeys + (pk & 0xffffu), srids + (pk & 0xffffu), pk >> 16);
		}

		// 4b. home
		for (uint32_t i = tid; i < size; i += THREADS) {
			st_stream_u64(dst_keys + i, skeys[i]);
			st_stream_u64(dst_rids + i, srids[i]);
		}
		__syncthreads();
#undef MSB64_BIN_SLOT
	}
}

} // namespace msb64
