// msb64_histogram.cuh -- per-segment digit histogram (replaces histogram(), msb_64.c:701-738).
//
// Algorithmic traffic: 8 bytes read per key, nothing written but the counters.
// One block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; every warp keeps a
// private NB-bin histogram in shared memory which is updated without atomics: the
// lanes of a warp first find their digit peers by ballots (match_digit) and only
// the lowest peer adds the peer count.  The warp histograms are summed and added
// to the segment's global counters whenever the block moves on to another segment.
#pragma once
#include "msb64_common.cuh"

namespace msb64 {

template <int BITS, int THREADS>
struct HistCfg {
	static constexpr int NB = 1 << BITS;
	static constexpr int WARPS = THREADS / 32;
	static constexpr int ITEMS = TILE / THREADS;
	static constexpr size_t SMEM = size_t(WARPS) * NB * sizeof(uint32_t);
};

template <int BITS, int THREADS>
__global__ void __launch_bounds__(THREADS)
histogram_kernel(const Ctx c, const int level, const int shift)
{
	using Cfg = HistCfg<BITS, THREADS>;
	constexpr int NB = Cfg::NB, WARPS = Cfg::WARPS, ITEMS = Cfg::ITEMS;
	static_assert(ITEMS % 2 == 0, "tile is loaded as 16-byte pairs");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint32_t *whist = reinterpret_cast<uint32_t *>(smem_raw);     // [WARPS][NB]

	const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
	const uint32_t ntiles = c.ctl->ntiles[level];
	const Seg *segs = ((level & 1) ? c.segs[1] : c.segs[0]);
	const Tile *tiles = ((level & 1) ? c.tiles[1] : c.tiles[0]);
	uint32_t *hist = ((level & 1) ? c.hist[1] : c.hist[0]);
	uint32_t *mine = whist + warp * NB;

	for (int i = tid; i < WARPS * NB; i += THREADS) whist[i] = 0;
	__syncthreads();

	uint32_t cur_seg = 0xffffffffu;
	for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const Tile tile = tiles[t];
		if (tile.seg != cur_seg) {
			if (cur_seg != 0xffffffffu) {
				__syncthreads();
				for (int b = tid; b < NB; b += THREADS) {
					uint32_t sum = 0;
#pragma unroll
					for (int w = 0; w < WARPS; ++w) {
						sum += whist[w * NB + b];
						whist[w * NB + b] = 0;
					}
					if (sum) atomicAdd(&hist[size_t(cur_seg) * NB + b], sum);
				}
				__syncthreads();
			}
			cur_seg = tile.seg;
		}
		const Seg s = segs[tile.seg];
		const uint64_t *keys = (s.buf ? c.keys[1] : c.keys[0]);
		const uint32_t end = s.begin + s.size;
		const uint32_t lo = seg_tile_origin(s.begin) + tile.idx * TILE;
		const bool full = lo >= s.begin && lo + TILE <= end;

		uint64_t k[ITEMS];
		if (full) {
#pragma unroll
			for (int j = 0; j < ITEMS / 2; ++j) {
				const ulonglong2 v = ld_stream_u64x2(keys + lo + (j * THREADS + tid) * 2);
				k[2 * j] = v.x;
				k[2 * j + 1] = v.y;
			}
#pragma unroll
			for (int j = 0; j < ITEMS; ++j) {
				const uint32_t d = uint32_t(k[j] >> shift) & (NB - 1);
				const uint32_t peers = match_digit<BITS>(d);
				if ((peers & lanemask_lt()) == 0) mine[d] += __popc(peers);
				__syncwarp();
			}
		} else {
#pragma unroll
			for (int j = 0; j < ITEMS; ++j) {
				const uint32_t e = lo + ((j >> 1) * THREADS + tid) * 2 + (j & 1);
				const bool valid = e >= s.begin && e < end;
				const uint64_t key = valid ? ld_stream_u64(keys + e) : 0;
				const uint32_t d = uint32_t(key >> shift) & (NB - 1);
				uint32_t peers = match_digit<BITS>(d);
				peers &= __ballot_sync(0xffffffffu, valid);
				if (valid && (peers & lanemask_lt()) == 0) mine[d] += __popc(peers);
				__syncwarp();
			}
		}
	}
	if (cur_seg != 0xffffffffu) {
		__syncthreads();
		for (int b = tid; b < NB; b += THREADS) {
			uint32_t sum = 0;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) sum += whist[w * NB + b];
			if (sum) atomicAdd(&hist[size_t(cur_seg) * NB + b], sum);
		}
	}
}

} // namespace msb64
