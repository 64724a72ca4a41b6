// msb64_local_sort.cuh -- finish of small buckets entirely in shared memory (the role of
// insertsort / combsort and the in-cache partition_ip levels, msb_64.c:126-149, 980-1005,
// 740-770).  Algorithmic traffic: 32 bytes per pair (16 read + 16 written), once.
//
// A unit is a run of neighbouring buckets of at most LOCAL_CAP pairs.  Its keys agree
// on some prefix and may differ anywhere below; the block
//   1. loads keys and rids (coalesced) and ORs together key ^ first_key: the set bits
//      are exactly the bit positions in which the unit's keys differ;
//   2. counting-sorts the pairs in shared memory on the top `b` differing bits
//      (b <= LOCAL_BITS): histogram, block scan, placement;
//   3. finishes every bin that still holds different keys: short bins by a serial
//      insertion sort of one thread (msb_64.c:126-149 does the same below 20 items),
//      long bins by a block-wide bitonic network (the fallback for adversarial bit
//      patterns; random keys never reach it);
//   4. writes the sorted unit to the caller's arrays, coalesced.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

constexpr int LOCAL_THREADS = 256;
constexpr int LOCAL_ITEMS = LOCAL_CAP / LOCAL_THREADS;
constexpr int LOCAL_BITS = 12;
constexpr uint32_t LOCAL_INSERT_MAX = 24;       // bins up to this size: one thread, insertion sort
constexpr size_t LOCAL_SMEM = size_t(LOCAL_CAP) * 16 + (size_t(1) << LOCAL_BITS) * 4
			      + (LOCAL_CAP / LOCAL_INSERT_MAX + 2) * 4 + 64 * 4 + 16 * 8;

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

__global__ void __launch_bounds__(LOCAL_THREADS)
local_sort_kernel(const Ctx c)
{
	constexpr int THREADS = LOCAL_THREADS, ITEMS = LOCAL_ITEMS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);        // [LOCAL_CAP]
	uint64_t *srids = skeys + LOCAL_CAP;                             // [LOCAL_CAP]
	uint32_t *bins = reinterpret_cast<uint32_t *>(srids + LOCAL_CAP);// [1 << LOCAL_BITS]
	uint32_t *big = bins + (1u << LOCAL_BITS);                       // long bins: base | size << 16
	uint32_t *scratch = big + (LOCAL_CAP / LOCAL_INSERT_MAX + 2);    // [64]
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
		const uint64_t first = __shfl_sync(0xffffffffu, k[0], 0);   // warp 0: element 0
		if (tid == 0) {
			wor[8] = first;
			s_nbig = 0;
		}
		__syncthreads();
		const uint64_t k0 = wor[8];
		uint64_t diff = 0;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (uint32_t(j * THREADS + tid) < size) diff |= k[j] ^ k0;
		uint32_t dlo = __reduce_or_sync(0xffffffffu, uint32_t(diff));
		uint32_t dhi = __reduce_or_sync(0xffffffffu, uint32_t(diff >> 32));
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

		// 2. counting sort on the top differing bits
		const int top = 63 - __clzll(diff);                      // highest differing bit
		int b = 32 - __clz(size - 1);                            // ~log2(size) bins ...
		b = min(max(b + 1, 5), LOCAL_BITS);                      // ... times two
		b = min(b, top + 1);
		const int shift = top + 1 - b;
		const uint32_t nb = 1u << b, dmask = nb - 1;

		for (uint32_t i = tid; i < nb; i += THREADS) bins[i] = 0;
		__syncthreads();
		uint32_t rank[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < size) rank[j] = atomicAdd(&bins[uint32_t(k[j] >> shift) & dmask], 1u);
		}
		__syncthreads();
		// exclusive scan over bins (thread handles nb/THREADS consecutive bins), pack base | count << 16
		{
			constexpr uint32_t MAXPER = (1u << LOCAL_BITS) / THREADS;
			const uint32_t per = (nb + THREADS - 1) / THREADS;
			uint32_t cnt[MAXPER], sum = 0;
#pragma unroll
			for (uint32_t q = 0; q < MAXPER; ++q) {
				const uint32_t bi = tid * per + q;
				cnt[q] = (q < per && bi < nb) ? bins[bi] : 0;
				sum += cnt[q];
			}
			uint32_t total;
			uint32_t base = block_exclusive_scan<THREADS>(sum, scratch, &total);
#pragma unroll
			for (uint32_t q = 0; q < MAXPER; ++q) {
				const uint32_t bi = tid * per + q;
				if (q < per && bi < nb) {
					bins[bi] = base | (cnt[q] << 16);
					base += cnt[q];
				}
			}
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < size) {
				const uint32_t p = (bins[uint32_t(k[j] >> shift) & dmask] & 0xffffu) + rank[j];
				skeys[p] = k[j];
				srids[p] = r[j];
			}
		}
		__syncthreads();

		// 3. bins that may still hold different keys
		if (shift > 0) {
			for (uint32_t bi = tid; bi < nb; bi += THREADS) {
				const uint32_t packed = bins[bi];
				const uint32_t base = packed & 0xffffu, cnt = packed >> 16;
				if (cnt < 2) continue;
				if (cnt <= LOCAL_INSERT_MAX) {
					for (uint32_t i = base + 1; i < base + cnt; ++i) {
						const uint64_t key = skeys[i];
						if (key >= skeys[i - 1]) continue;
						const uint64_t rid = srids[i];
						uint32_t j = i;
						do {
							skeys[j] = skeys[j - 1];
							srids[j] = srids[j - 1];
							--j;
						} while (j > base && key < skeys[j - 1]);
						skeys[j] = key;
						srids[j] = rid;
					}
				} else {
					big[atomicAdd(&s_nbig, 1u)] = packed;
				}
			}
			__syncthreads();
			const uint32_t nbig = s_nbig;
			for (uint32_t q = 0; q < nbig; ++q) {
				const uint32_t packed = big[q];
				block_bitonic(skeys + (packed & 0xffffu), srids + (packed & 0xffffu), packed >> 16);
			}
		}

		// 4. home
		for (uint32_t i = tid; i < size; i += THREADS) {
			st_stream_u64(dst_keys + i, skeys[i]);
			st_stream_u64(dst_rids + i, srids[i]);
		}
		__syncthreads();
	}
}

} // namespace msb64
