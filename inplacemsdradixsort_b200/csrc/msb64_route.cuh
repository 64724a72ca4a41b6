// msb64_route.cuh -- the two device steps of the multi-GPU range partition (the role of
// the reference's sample / range_histogram / partition-to-blocks phase across NUMA nodes,
// msb_64.c:239-351, 497-699, 1546-1606):
//
//   digit_histogram_kernel  counts of the top `bits` key bits of a rank's slice; the
//                           per-rank histograms are all-gathered and cut into one
//                           contiguous bin range per rank by the host (distributed.py);
//   route_kernel            groups the slice by destination rank (bin -> rank table).  Every
//                           destination has its own output arrays: either slices of one
//                           local send buffer (every rank's share is then one contiguous
//                           run to hand to ncclSend), or -- the fused compute + exchange
//                           form -- the receive buffers of the peer GPUs themselves, mapped
//                           through CUDA IPC: the kernel's coalesced stores travel over
//                           NVLink / NVSwitch straight into the destination's HBM and no
//                           separate exchange pass exists.  With <= 64 destinations the
//                           runs written per tile are long (>= 64 pairs, 512 bytes), the
//                           regime in which both the HBM write path (tools/permcopy.cu)
//                           and NVLink run at full rate.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

constexpr int ROUTE_THREADS = 256;
constexpr int ROUTE_ITEMS = TILE / ROUTE_THREADS;
constexpr int ROUTE_MAX_DEST = 64;
constexpr int ROUTE_MAX_BITS = 12;

__global__ void __launch_bounds__(256)
digit_histogram_kernel(const uint64_t *keys, uint64_t n, int shift, int bits,
		       unsigned long long *hist)
{
	extern __shared__ uint32_t sh[];
	const uint32_t nb = 1u << bits;
	const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
	const uint64_t lo = per_block * blockIdx.x;
	const uint64_t hi = lo + per_block < n ? lo + per_block : n;
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	// 32-bit block counters: a block never sees more than 2^32 keys (n <= MSB64_MAX_PAIRS)
	for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
		atomicAdd(&sh[uint32_t(ld_stream_u64(keys + i) >> shift) & (nb - 1)], 1u);
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
		if (sh[i]) atomicAdd(&hist[i], (unsigned long long) sh[i]);
}

// Output arrays of every destination (kernel parameter, 1 KiB).
struct RouteDst {
	uint64_t *keys[ROUTE_MAX_DEST];
	uint64_t *rids[ROUTE_MAX_DEST];
};

// cursors[d] = next free slot of this source in destination d's output arrays (initialised
// by the host: exclusive prefix of the send counts for a local send buffer, number of pairs
// the lower-ranked sources send to d for a peer's receive buffer).
__global__ void __launch_bounds__(ROUTE_THREADS, 2)
route_kernel(const uint64_t *keys, const uint64_t *rids, uint32_t n, int shift, int bits,
	     const uint8_t *bin_to_dest, int ndest, uint32_t *cursors, const RouteDst dst)
{
	constexpr int THREADS = ROUTE_THREADS, ITEMS = ROUTE_ITEMS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *skeys = reinterpret_cast<uint64_t *>(smem_raw);             // [TILE]
	uint64_t *srids = skeys + TILE;                                       // [TILE]
	uint32_t *cnt = reinterpret_cast<uint32_t *>(srids + TILE);           // [ROUTE_MAX_DEST + 32]
	uint32_t *lbase = cnt + ROUTE_MAX_DEST + 32;                          // [ROUTE_MAX_DEST]
	uint32_t *delta = lbase + ROUTE_MAX_DEST;                             // [ROUTE_MAX_DEST]
	uint64_t **okeys = reinterpret_cast<uint64_t **>(delta + ROUTE_MAX_DEST);   // [ROUTE_MAX_DEST]
	uint64_t **orids = okeys + ROUTE_MAX_DEST;                             // [ROUTE_MAX_DEST]
	uint8_t *table = reinterpret_cast<uint8_t *>(orids + ROUTE_MAX_DEST); // [1 << bits]

	const uint32_t tid = threadIdx.x, lane = lane_id();
	const uint32_t nb = 1u << bits;
	for (uint32_t i = tid; i < nb; i += THREADS) table[i] = bin_to_dest[i];
	if (tid < uint32_t(ndest)) {
		okeys[tid] = dst.keys[tid];
		orids[tid] = dst.rids[tid];
	}
	const uint32_t ntiles = (n + TILE - 1) / TILE;

	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const uint32_t lo = t * TILE;
		const uint32_t count = min(TILE, n - lo);
		for (uint32_t i = tid; i < ROUTE_MAX_DEST + 32; i += THREADS) cnt[i] = 0;
		__syncthreads();

		uint64_t k[ITEMS], r[ITEMS];
		uint32_t rank[ITEMS], dest[ITEMS];
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			k[j] = i < count ? ld_stream_u64(keys + lo + i) : 0;
		}
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			r[j] = i < count ? ld_stream_u64(rids + lo + i) : 0;
		}
		// branch-free ranking (see tile_ranks in msb64_scatter.cuh); slots past the end of
		// the slice count into per-lane dummy bins
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			dest[j] = i < count ? uint32_t(table[uint32_t(k[j] >> shift) & (nb - 1)])
					    : uint32_t(ROUTE_MAX_DEST) + lane;
			rank[j] = atomicAdd(&cnt[dest[j]], 1u);
		}
		__syncthreads();
		if (tid < uint32_t(ndest)) {
			// at most 64 destinations: every owner thread sums its predecessors
			uint32_t before = 0;
			for (uint32_t d = 0; d < tid; ++d) before += cnt[d];
			const uint32_t c = cnt[tid];
			const uint32_t g = c ? atomicAdd(&cursors[tid], c) : 0;
			lbase[tid] = before;
			delta[tid] = g - before;
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t i = j * THREADS + tid;
			if (i < count) {
				const uint32_t p = lbase[dest[j]] + rank[j];
				skeys[p] = k[j];
				srids[p] = r[j];
			}
		}
		__syncthreads();
		for (uint32_t i = tid; i < count; i += THREADS) {
			const uint64_t key = skeys[i];
			const uint32_t d = table[uint32_t(key >> shift) & (nb - 1)];
			const uint32_t at = delta[d] + i;
			st_stream_u64(okeys[d] + at, key);
			st_stream_u64(orids[d] + at, srids[i]);
		}
		__syncthreads();
	}
}

constexpr size_t route_smem(int bits)
{
	return size_t(TILE) * 16 + (ROUTE_MAX_DEST + 32 + 2 * ROUTE_MAX_DEST) * 4 + 2 * ROUTE_MAX_DEST * 8
	       + (size_t(1) << bits);
}

} // namespace msb64
