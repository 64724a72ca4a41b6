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
#include "msb64_scatter.cuh"   // bulk-copy / mbarrier primitives, tile_ranks

namespace msb64 {

constexpr int ROUTE_THREADS = 256;
constexpr int ROUTE_ITEMS = TILE / ROUTE_THREADS;
constexpr int ROUTE_MAX_DEST = 64;
constexpr int ROUTE_MAX_BITS = 13;

// digit = ((key >> shift) - origin) & (2^bits - 1); minmax (optional): [0] = smallest, [1] =
// largest key seen (initialised by the host to ~0 and 0).
__global__ void __launch_bounds__(256)
digit_histogram_kernel(const uint64_t *keys, uint64_t n, int shift, int bits, uint32_t origin,
		       unsigned long long *hist, unsigned long long *minmax)
{
	extern __shared__ uint32_t sh[];
	const uint32_t nb = 1u << bits;
	const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
	const uint64_t lo = per_block * blockIdx.x;
	const uint64_t hi = lo + per_block < n ? lo + per_block : n;
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	// 32-bit block counters: a block never sees more than 2^32 keys (n <= MSB64_MAX_PAIRS)
	unsigned long long kmin = ~0ull, kmax = 0;
	// four independent loads per thread and trip: enough bytes in flight to fill HBM
	for (uint64_t i = lo + threadIdx.x; i < hi; i += 4ull * blockDim.x) {
		uint64_t k[4];
		bool ok[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const uint64_t at = i + uint64_t(u) * blockDim.x;
			ok[u] = at < hi;
			k[u] = ok[u] ? ld_stream_u64(keys + at) : 0;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (ok[u]) {
				kmin = k[u] < kmin ? k[u] : kmin;
				kmax = k[u] > kmax ? k[u] : kmax;
				atomicAdd(&sh[(uint32_t(k[u] >> shift) - origin) & (nb - 1)], 1u);
			}
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
		if (sh[i]) atomicAdd(&hist[i], (unsigned long long) sh[i]);
	if (minmax) {
		for (int d = 16; d; d >>= 1) {
			const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, d);
			const unsigned long long b = __shfl_xor_sync(0xffffffffu, kmax, d);
			kmin = a < kmin ? a : kmin;
			kmax = b > kmax ? b : kmax;
		}
		if (lane_id() == 0 && kmin <= kmax) {
			atomicMin(&minmax[0], kmin);
			atomicMax(&minmax[1], kmax);
		}
	}
}

// Output arrays of every destination (kernel parameter, 1 KiB).
struct RouteDst {
	uint64_t *keys[ROUTE_MAX_DEST];
	uint64_t *rids[ROUTE_MAX_DEST];
};

constexpr uint32_t ROUTE_ALIGN = 16;     // pairs: every destination's run is written in 128-byte aligned pieces

template <int ND>
struct RouteCfg {
	// padded destination-ordered positions: every non-empty destination adds up to ROUTE_ALIGN - 1
	// slots in front of its run and as many behind it
	static constexpr uint32_t SLOTS = TILE + 2 * ROUTE_ALIGN * ND;
	static constexpr size_t SMEM = size_t(TILE) * 16                        // keys + rids of the tile (bulk-copied)
				       + size_t(SLOTS) * 2                      // source slot by padded position
				       + (ND + 32) * 4                          // per-destination counters (+ dummies)
				       + 3 * ND * 4                             // padded size, local base, global - local base
				       + 2 * ND * 8                             // output pointers
				       + 16;                                    // mbarrier
};

// cursors[d] = next free slot of this source in destination d's output arrays (initialised
// by the host: exclusive prefix of the send counts for a local send buffer, number of pairs
// the lower-ranked sources send to d for a peer's receive buffer).
//
// Same structure as scatter_kernel (msb64_scatter.cuh): the tile lands in shared memory by
// bulk asynchronous copies, the pairs stay where they landed, a 2-byte source slot per pair
// is written in destination order and the write-out gathers through it; three blocks per SM
// (ND = 16) keep enough stores in flight for NVLink (tools/p2p_bench.cu: plain coalesced
// 8-byte stores reach 0.69 TB/s per direction with >= 4 x 256 threads per SM, copy engines
// 0.78 TB/s).  Stores to a peer are not merged by any cache on the way, so the write-out is
// laid out in GLOBAL 128-byte lines: the destination-ordered positions are padded so that
// position p of destination d goes to element delta[d] + p with delta[d] a multiple of 16 --
// every warp store is two whole lines instead of a line and two fragments.
// The bin -> destination table (4 KiB at 12 bits) is read through L1.
template <int ND>
__global__ void __launch_bounds__(ROUTE_THREADS, ND <= 16 ? 3 : 2)
route_kernel(const uint64_t *keys, const uint64_t *rids, uint32_t n, int shift, int bits, uint32_t origin,
	     const uint8_t *__restrict__ bin_to_dest, int ndest, uint32_t *cursors, const RouteDst dst)
{
	constexpr int THREADS = ROUTE_THREADS, ITEMS = ROUTE_ITEMS;
	using Cfg = RouteCfg<ND>;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *kin = reinterpret_cast<uint64_t *>(smem_raw);               // [TILE]
	uint64_t *rin = kin + TILE;                                           // [TILE]
	uint16_t *sidx = reinterpret_cast<uint16_t *>(rin + TILE);            // [SLOTS]
	uint32_t *cnt = reinterpret_cast<uint32_t *>(sidx + Cfg::SLOTS);      // [ND + 32]
	uint32_t *psize = cnt + ND + 32;                                      // [ND] padded run length
	uint32_t *lbase = psize + ND;                                         // [ND] first position of the run
	uint32_t *delta = lbase + ND;                                         // [ND] global element - position
	uint64_t **okeys = reinterpret_cast<uint64_t **>(delta + ND);         // [ND]
	uint64_t **orids = okeys + ND;                                        // [ND]
	uint64_t *bar = reinterpret_cast<uint64_t *>(orids + ND);
	__shared__ uint32_t s_slots;

	const uint32_t tid = threadIdx.x, lane = lane_id();
	const uint32_t dmask = (1u << bits) - 1;
	const uint32_t ntiles = (n + TILE - 1) / TILE;
	if (blockIdx.x >= ntiles) return;
	// thread 0: a tile that lies inside the array completely is fetched by bulk copies
	auto start_copy = [&](uint32_t t) {
		if (t >= ntiles || (t + 1) * uint64_t(TILE) > n) return;
		mbar_expect_tx(bar, TILE * 16);
		bulk_copy_g2s(kin, keys + size_t(t) * TILE, TILE * 8, bar);
		bulk_copy_g2s(rin, rids + size_t(t) * TILE, TILE * 8, bar);
	};
	if (tid == 0) {
		mbar_init(bar, 1);
		start_copy(blockIdx.x);
	}
	if (tid < uint32_t(ndest)) {
		okeys[tid] = dst.keys[tid];
		orids[tid] = dst.rids[tid];
	}
	for (uint32_t i = tid; i < ND + 32; i += THREADS) cnt[i] = 0;
	__syncthreads();

	uint32_t parity = 0;
	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const uint32_t lo = t * TILE;
		const uint32_t count = min(TILE, n - lo);
		if (count == TILE) {
			mbar_wait(bar, parity);
			parity ^= 1u;
		} else {
			// the array's tail: plain loads
			for (uint32_t i = tid; i < count; i += THREADS) {
				kin[i] = ld_stream_u64(keys + lo + i);
				rin[i] = ld_stream_u64(rids + lo + i);
			}
			__syncthreads();
		}
		// destination of every pair, rank among the tile's pairs with the same destination
		// (slots past the end count into per-lane dummy counters)
		uint32_t dr[ITEMS];
		{
			const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(kin);
#pragma unroll
			for (int jj = 0; jj < ITEMS / 2; ++jj) {
				const ulonglong2 v = k2[jj * THREADS + tid];
				const uint32_t s0 = (jj * THREADS + tid) * 2;
				dr[2 * jj] = s0 < count ? uint32_t(__ldg(bin_to_dest + ((uint32_t(v.x >> shift) - origin) & dmask))) : ND + lane;
				dr[2 * jj + 1] = s0 + 1 < count ? uint32_t(__ldg(bin_to_dest + ((uint32_t(v.y >> shift) - origin) & dmask))) : ND + lane;
			}
		}
		tile_ranks<ITEMS, ND>(cnt, dr);
		__syncthreads();
		// every destination's owner thread claims the run's slice of the output and pads
		// the run to the output's 128-byte lines
		uint32_t g = 0, c = 0, head = 0;
		if (tid < uint32_t(ndest)) {
			c = cnt[tid];
			g = c ? atomicAdd(&cursors[tid], c) : 0;
			head = c ? (g & (ROUTE_ALIGN - 1)) : 0;
			psize[tid] = (head + c + ROUTE_ALIGN - 1) & ~(ROUTE_ALIGN - 1);
		}
		__syncthreads();
		if (tid < uint32_t(ndest)) {
			uint32_t before = 0;
			for (uint32_t d = 0; d < tid; ++d) before += psize[d];
			const uint32_t first = before + head;
			lbase[tid] = first;
			delta[tid] = g - first;
			for (uint32_t p = before; p < first; ++p) sidx[p] = 0xffffu;            // padding in front
			for (uint32_t p = first + c; p < before + psize[tid]; ++p) sidx[p] = 0xffffu;   // and behind
			if (tid + 1 == uint32_t(ndest)) s_slots = before + psize[tid];
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t d = dr[j] >> RANK_BITS;
			if (d < ND) sidx[lbase[d] + (dr[j] & RANK_MASK)] = uint16_t(((j >> 1) * THREADS + tid) * 2 + (j & 1));
		}
		__syncthreads();
		const uint32_t slots = s_slots;
#pragma unroll 4
		for (uint32_t p = tid; p < slots; p += THREADS) {
			const uint32_t s = sidx[p];
			if (s != 0xffffu) {
				const uint64_t key = kin[s];
				const uint64_t rid = rin[s];
				const uint32_t d = __ldg(bin_to_dest + ((uint32_t(key >> shift) - origin) & dmask));
				const uint32_t at = delta[d] + p;
				st_stream_u64(okeys[d] + at, key);
				st_stream_u64(orids[d] + at, rid);
			}
		}
		__syncthreads();                                      // every thread is done with sidx / delta / the tile
		for (uint32_t i = tid; i < ND + 32; i += THREADS) cnt[i] = 0;
		if (tid == 0) start_copy(t + gridDim.x);
		__syncthreads();
	}
}


// ---------------------------------------------------------------------------------------
// bucket_route_kernel -- first pass of the pipelined multi-GPU sort (msb64_shard.cuh): the
// rank's pairs are grouped by BUCKET = table[digit], where a bucket is one (destination GPU,
// sub-range) cell of the range partition: destination d owns buckets [d * subs, (d + 1) * subs)
// and inside a destination the sub-ranges are ascending key ranges of near-equal count.  It is
// an MSD partition pass on splitters instead of a bit field (the reference partitions on
// sampled delimiters the same way, msb_64.c:497-699), written locally at HBM speed: the
// buckets of other GPUs land bucket by bucket in a staging buffer, from where whole buckets
// travel over NVLink as large contiguous copies while the destination already sorts the
// buckets that have arrived; the rank's own buckets go straight to their final place in its
// receive buffer.
//
// Structure of scatter_kernel (msb64_scatter.cuh) on a flat array: bulk-copied tile, shared
// atomics for the ranks, one global atomicAdd per non-empty bucket on cursors[bucket], a
// 2-byte source-slot permutation and a coalesced write-out.  Algorithmic traffic 32 B/pair.
struct BucketOut {
	// [d] = receive buffer of rank d (this process' mapping of it; [self] = the rank's own),
	// [ROUTE_MAX_DEST] = the local staging buffer
	uint64_t *keys[ROUTE_MAX_DEST + 1];
	uint64_t *rids[ROUTE_MAX_DEST + 1];
	uint32_t subs;       // buckets per destination
	uint32_t ndirect;    // sub-ranges [0, ndirect) of every destination are stored straight into its receive
			     // buffer, over NVLink for the peers: the fused compute + exchange part of the pass
	uint32_t self;       // all buckets of this rank go straight to its own receive buffer
};

#ifndef MSB64_BUCKET_THREADS
#define MSB64_BUCKET_THREADS 256
#endif
#ifndef MSB64_BUCKET_MINB
#define MSB64_BUCKET_MINB 3
#endif
template <int NBK>
struct BucketCfg {
	static constexpr int THREADS = MSB64_BUCKET_THREADS;   // 256 x 3 blocks per SM: 5.58 ms per 2^30 pairs (512 x 2: 6.12 ms -- the opposite of the scatter pass)
	static constexpr int MINB = MSB64_BUCKET_MINB;
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr size_t SMEM = size_t(TILE) * 16 + size_t(TILE) * 2 + size_t(NBK + 32) * 4 + size_t(NBK) * 4
				       + 64 * 4 + 16
				       + size_t(ROUTE_MAX_DEST + 1) * 16 + 2 * NBK;    // output pointers by target, target of every bucket, list of the peers' buckets
};

template <int NBK>
__global__ void __launch_bounds__(BucketCfg<NBK>::THREADS, BucketCfg<NBK>::MINB)
bucket_route_kernel(const uint64_t *keys, const uint64_t *rids, uint32_t n, int shift, uint32_t tmask,
		    uint32_t origin, const uint8_t *__restrict__ table, uint32_t *cursors, const BucketOut out)
{
	using Cfg = BucketCfg<NBK>;
	constexpr int THREADS = Cfg::THREADS, ITEMS = Cfg::ITEMS;
	static_assert(NBK <= THREADS, "one owner thread per bucket");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *kin = reinterpret_cast<uint64_t *>(smem_raw);               // [TILE]
	uint64_t *rin = kin + TILE;                                           // [TILE]
	uint16_t *sidx = reinterpret_cast<uint16_t *>(rin + TILE);            // [TILE]
	uint32_t *cnt = reinterpret_cast<uint32_t *>(sidx + TILE);            // [NBK + 32]
	uint32_t *delta = cnt + NBK + 32;                                     // [NBK]
	uint32_t *scratch = delta + NBK;                                      // [64]
	uint64_t *bar = reinterpret_cast<uint64_t *>(scratch + 64);           // [2] (one used)
	uint64_t **okeys = reinterpret_cast<uint64_t **>(bar + 2);            // [ROUTE_MAX_DEST + 1]
	uint64_t **orids = okeys + ROUTE_MAX_DEST + 1;                        // [ROUTE_MAX_DEST + 1]
	uint8_t *target = reinterpret_cast<uint8_t *>(orids + ROUTE_MAX_DEST + 1);   // [NBK] index into okeys / orids
	uint8_t *plist = target + NBK;                                        // [NBK] buckets stored into a peer's memory
	__shared__ uint32_t s_npeer;

	const uint32_t tid = threadIdx.x, lane = lane_id();
	const uint32_t ntiles = (n + TILE - 1) / TILE;
	if (blockIdx.x >= ntiles) return;
	for (uint32_t i = tid; i <= ROUTE_MAX_DEST; i += THREADS) {
		okeys[i] = out.keys[i];
		orids[i] = out.rids[i];
	}
	for (uint32_t b = tid; b < NBK; b += THREADS) {
		const uint32_t d = b / out.subs, sub = b - d * out.subs;
		target[b] = uint8_t(d < ROUTE_MAX_DEST && (d == out.self || sub < out.ndirect) ? d : ROUTE_MAX_DEST);
	}
	if (tid == 0) {
		uint32_t np = 0;
		for (uint32_t b = 0; b < NBK; ++b) {
			const uint32_t d = b / out.subs, sub = b - d * out.subs;
			if (d < ROUTE_MAX_DEST && d != out.self && sub < out.ndirect) plist[np++] = uint8_t(b);
		}
		s_npeer = np;
	}
	auto start_copy = [&](uint32_t t) {
		if (t >= ntiles || (t + 1) * uint64_t(TILE) > n) return;
		mbar_expect_tx(bar, TILE * 16);
		bulk_copy_g2s(kin, keys + size_t(t) * TILE, TILE * 8, bar);
		bulk_copy_g2s(rin, rids + size_t(t) * TILE, TILE * 8, bar);
	};
	auto bucket_of = [&](uint64_t key) -> uint32_t {
		return uint32_t(__ldg(table + ((uint32_t(key >> shift) - origin) & tmask)));
	};
	if (tid == 0) {
		mbar_init(bar, 1);
		start_copy(blockIdx.x);
	}
	for (uint32_t i = tid; i < NBK + 32; i += THREADS) cnt[i] = 0;
	__syncthreads();

	uint32_t parity = 0;
	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const uint32_t lo = t * TILE;
		const uint32_t count = min(TILE, n - lo);
		if (count == TILE) {
			mbar_wait(bar, parity);
			parity ^= 1u;
		} else {
			for (uint32_t i = tid; i < count; i += THREADS) {
				kin[i] = ld_stream_u64(keys + lo + i);
				rin[i] = ld_stream_u64(rids + lo + i);
			}
			__syncthreads();
		}
		uint32_t dr[ITEMS];
		{
			const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(kin);
#pragma unroll
			for (int jj = 0; jj < ITEMS / 2; ++jj) {
				const ulonglong2 v = k2[jj * THREADS + tid];
				const uint32_t s0 = (jj * THREADS + tid) * 2;
				dr[2 * jj] = s0 < count ? bucket_of(v.x) : NBK + lane;
				dr[2 * jj + 1] = s0 + 1 < count ? bucket_of(v.y) : NBK + lane;
			}
		}
		tile_ranks<ITEMS, NBK>(cnt, dr);
		__syncthreads();
		// owner thread of every bucket: claim the tile's slice, exclusive scan over the buckets
		const uint32_t tot = tid < NBK ? cnt[tid] : 0;
		const uint32_t g = tot ? atomicAdd(&cursors[tid], tot) : 0;
		uint32_t total;
		const uint32_t lbase = block_exclusive_scan<THREADS>(tot, scratch, &total);
		if (tid < NBK) {
			cnt[tid] = lbase;
			delta[tid] = g - lbase;
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < ITEMS; ++j) {
			const uint32_t d = dr[j] >> RANK_BITS;
			if (d < NBK) sidx[cnt[d] + (dr[j] & RANK_MASK)] = uint16_t(((j >> 1) * THREADS + tid) * 2 + (j & 1));
		}
		__syncthreads();
		// write-out.  Local targets (staging, the rank's own receive buffer): position i of the
		// bucket-ordered tile goes to delta[bucket] + i, consecutive lanes to consecutive addresses.
#pragma unroll 8
		for (uint32_t i = tid; i < count; i += THREADS) {
			const uint32_t s = sidx[i];
			const uint64_t key = kin[s];
			const uint32_t b = bucket_of(key);
			const uint32_t to = target[b];
			if (to == ROUTE_MAX_DEST || to == out.self) {
				const uint32_t at = delta[b] + i;
				st_stream_u64(okeys[to] + at, key);
				st_stream_u64(orids[to] + at, rin[s]);
			}
		}
		// Buckets that live in a peer's memory: nothing merges stores on their way over NVLink,
		// so a warp takes a bucket's run and writes it in the DESTINATION's 256-byte windows
		// (lane = element index mod 32): every store instruction is at most two 128-byte lines,
		// whole ones except at the run's two ends.
		{
			const uint32_t npeer = s_npeer;
			for (uint32_t q = tid >> 5; q < npeer; q += THREADS / 32) {
				const uint32_t b = plist[q];
				const uint32_t first = cnt[b];
				const uint32_t len = (b + 1 < NBK ? cnt[b + 1] : total) - first;
				if (!len) continue;
				const uint32_t dl = delta[b], g0 = dl + first, g1 = g0 + len;
				uint64_t *pk = okeys[target[b]], *pr = orids[target[b]];
				for (uint32_t e = (g0 & ~31u) + lane; e < g1; e += 32)
					if (e >= g0) {
						const uint32_t s = sidx[e - dl];
						st_stream_u64(pk + e, kin[s]);
						st_stream_u64(pr + e, rin[s]);
					}
			}
		}
		__syncthreads();
		for (uint32_t i = tid; i < NBK + 32; i += THREADS) cnt[i] = 0;
		if (tid == 0) start_copy(t + gridDim.x);
		__syncthreads();
	}
}

// ---- completion flags of the pipelined exchange (msb64_shard.cuh)
// signal: one thread per peer writes `epoch` into that peer's flag word for (lane, sub, this
// source); the copies that precede it in the stream have completed, the fence orders them
// before the flag for an observer on another GPU.
struct FlagDst {
	uint32_t *flag[ROUTE_MAX_DEST];
};
__global__ void shard_signal_kernel(const FlagDst dst, int world, int self, uint32_t slot, uint32_t epoch)
{
	const int d = threadIdx.x;
	if (d >= world || d == self) return;
	__threadfence_system();
	*reinterpret_cast<volatile uint32_t *>(dst.flag[d] + slot) = epoch;
}
// the same for the sub-ranges the route pass stored straight into the peers' memory: thread
// (peer, lane, sub-range) raises that flag word; runs behind the route kernel in its stream
__global__ void shard_signal_direct_kernel(const FlagDst dst, int world, int self, int subs, int ndirect, uint32_t epoch)
{
	__threadfence_system();
	for (int i = threadIdx.x; i < world * 2 * ndirect; i += blockDim.x) {
		const int d = i / (2 * ndirect), r = i - d * 2 * ndirect, lane = r / ndirect, sub = r - lane * ndirect;
		if (d != self)
			*reinterpret_cast<volatile uint32_t *>(dst.flag[d] + (size_t(lane) * subs + sub) * world + self) = epoch;
	}
}
// wait: spins until every source's flag words of the given lanes of sub-range `sub` carry this epoch
__global__ void shard_wait_kernel(const uint32_t *flags, int world, int self, int subs, int sub, int lanes,
				  uint32_t epoch)
{
	const int t = threadIdx.x;
	if (t < world * lanes) {
		const int src = t % world, lane = t / world;
		if (src != self) {
			const volatile uint32_t *f = flags + (size_t(lane) * subs + sub) * world + src;
			while (int32_t(*f - epoch) < 0) __nanosleep(200);
		}
	}
	__syncthreads();
	__threadfence_system();
}

} // namespace msb64
