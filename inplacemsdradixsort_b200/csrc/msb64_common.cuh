// msb64_common.cuh -- device-side data model shared by the msb64 kernels (sm_100a only).
//
// The sort is an MSD radix sort over 64-bit keys with 64-bit rids riding along
// (reference: local_radixsort, msb_64.c:1007-1035).  One "level" consumes one
// digit.  The state between kernels lives in HBM:
//
//   buffers   A = the caller's arrays (buf 0), B = scratch of the same size (buf 1)
//   Seg       a bucket still too large for shared memory: [begin, begin+size) in
//             buffer `buf`; it is histogrammed and scattered at the next level
//   Tile      TILE consecutive element slots of one Seg; the unit of work of the
//             histogram and scatter kernels
//   Unit      a run of neighbouring small buckets (<= LOCAL_CAP pairs together)
//             finished by the local sort kernel entirely in shared memory
//   CopyTile  a piece of a finished bucket that ended in B and is copied to A
//
// Every list is filled by the plan kernel of the level above through atomic
// counters in Control; no kernel waits on another block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msb64 {

constexpr int MAX_LEVELS = 16;          // digits per key (64 bits / >= 4 bits)
constexpr int MAX_BITS = 11;            // widest digit a level may use
#ifndef MSB64_LOCAL_CAP
#define MSB64_LOCAL_CAP 4096
#endif
constexpr uint32_t LOCAL_CAP = MSB64_LOCAL_CAP;   // pairs the local sort holds in shared memory
// Largest unit the plan kernel files: two slots short of LOCAL_CAP, so that a unit that begins at
// an odd element still fits the local sort's 16-byte aligned window of LOCAL_CAP slots.
constexpr uint32_t UNIT_CAP = LOCAL_CAP - 2;
#ifndef MSB64_TILE
#define MSB64_TILE 4096
#endif
constexpr uint32_t TILE = MSB64_TILE;   // element slots per histogram/scatter tile
constexpr uint32_t COPY_TILE = 8192;    // pairs per copy tile
#ifndef MSB64_FUSE_MAX_BITS
#define MSB64_FUSE_MAX_BITS 14
#endif
constexpr int FUSE_MAX_BITS = MSB64_FUSE_MAX_BITS;   // level 0 + level 1 digit bits the fused histogram pass handles (4 << bits bytes of shared counters)

struct Seg {
	uint32_t begin;   // first element
	uint32_t size;    // > LOCAL_CAP by construction
	uint32_t buf;     // buffer holding it now (0 = A, 1 = B)
	uint32_t flags;   // SEG_* bits | position of the segment's digit << 8
};
// SEG_SKIP: every key has the same digit, nothing to move (set by this level's plan);
// SEG_HIST_READY: the histogram was computed by the level above (fused pass);
// SEG_WANT_BITS: the histogram pass also accumulates OR / AND of the segment's keys (asked
// for by the plan kernel when the level above found the segment's digit degenerate).
constexpr uint32_t SEG_SKIP = 1u, SEG_HIST_READY = 2u, SEG_WANT_BITS = 4u;
// The digit a segment is partitioned on is (key >> shift) & (2^bits - 1): the width comes from
// the level (schedule), the POSITION is the segment's own -- a segment whose keys agree on
// more bits than the schedule assumes (low-entropy keys) is moved down to its highest
// differing bit by the plan kernel instead of paying one histogram pass per dead digit.
__host__ __device__ inline uint32_t seg_flags(int shift, uint32_t f) { return (uint32_t(shift) << 8) | f; }
__host__ __device__ inline int seg_shift(uint32_t flags) { return int((flags >> 8) & 63u); }

// OR and AND of a segment's keys, accumulated by the histogram pass of SEG_WANT_BITS segments:
// OR & ~AND = the bits in which its keys differ.
struct SegBits {
	unsigned long long vor, vand;
};

struct Tile {
	uint32_t seg;     // index into this level's Seg list
	uint32_t idx;     // tile number inside the Seg
};

struct Unit {
	uint32_t begin;
	uint32_t size;    // 1..LOCAL_CAP
	uint32_t buf;     // where the pairs are now; they are always written to A
	uint32_t origin;  // unit_origin(): digit position and first digit of the run of buckets
};

// A unit is a run of buckets with consecutive digits first, first+1, ... of a `bits`-wide
// digit at bit position `shift`: its keys lie in a contiguous key range that starts at
// prefix | first << shift and agree on every bit above shift + bits.  The local sort
// subtracts that origin so that the keys of a unit made of several buckets spread evenly
// over its counting bins.  origin word: shift [0,6) | bits [6,10) | first digit [10,22) |
// bit 22: filed at level 0 (digits relative to the sort's key_lo, see msb64_b200.cu);
// bits = 0: nothing is known about the unit's keys (a whole small array).
constexpr uint32_t UNIT_LEVEL0 = 1u << 22;
__host__ __device__ inline uint32_t unit_origin(int shift, uint32_t first_digit, int bits, bool level0)
{
	return uint32_t(shift) | (uint32_t(bits) << 6) | (first_digit << 10) | (level0 ? UNIT_LEVEL0 : 0u);
}
__host__ __device__ inline uint64_t unit_origin_key(uint32_t origin)
{
	return uint64_t((origin >> 10) & 0xfffu) << (origin & 63u);
}

struct CopyTile {
	uint32_t begin;
	uint32_t size;    // 1..COPY_TILE, B -> A
};

struct Control {
	uint32_t nsegs[MAX_LEVELS + 1];
	uint32_t ntiles[MAX_LEVELS + 1];
	uint32_t nunits;         // units of the packed local sort, filed from the front of `units`
	uint32_t nslow;          // units of the general local sort, filed from the back
	uint32_t ncopies;
	uint32_t error;          // bit 0: Seg list overflow, 1: Tile, 2: Unit, 3: CopyTile
	uint32_t degenerate;     // segments whose scatter was skipped (statistics)
	uint32_t moved[MAX_LEVELS];   // pairs the scatter of each level moves (statistics)
	uint32_t local_pairs;    // pairs finished by the local sort (statistics)
	uint32_t nready[MAX_LEVELS + 1];  // segments of each level whose histogram the level above computed
	unsigned long long hist_keys;     // keys read by histogram passes (statistics)
};

// Everything a kernel needs, passed by value.
struct Ctx {
	uint64_t *keys[2];
	uint64_t *rids[2];
	Seg *segs[2];            // ping-pong by level parity
	Tile *tiles[2];
	uint32_t *hist[2];       // [seg][bin] counts, turned into write cursors by plan
	SegBits *segbits[2];     // OR / AND of every segment's keys
	Unit *units;
	CopyTile *copies;
	Control *ctl;
	uint32_t *fused;         // [2^bits0][2^bits1] level-1 digit counts per level-0 bin (fused histogram pass)
	uint32_t *status;        // page-locked host word: receives a non-zero Control.error at the end of the sort
	uint32_t begin;          // the sort's pairs are elements [begin, begin + n) of keys[] / rids[]
	uint32_t n;
	uint32_t end;            // begin + n: no kernel touches an element at or past it (rounded up to even)
	uint32_t max_segs, max_tiles, max_units, max_copies;
};

// Tiles of a Seg start at an even element so that 16-byte loads stay aligned.
__host__ __device__ inline uint32_t seg_tile_origin(uint32_t begin) { return begin & ~1u; }
__host__ __device__ inline uint32_t seg_tile_count(uint32_t begin, uint32_t size)
{
	return (begin + size - seg_tile_origin(begin) + TILE - 1) / TILE;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
// Streaming 16-byte / 8-byte global accesses: the sort touches every byte once per
// pass, so nothing is worth keeping in L1.
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const uint64_t *p)
{
	ulonglong2 v;
	asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];"
		     : "=l"(v.x), "=l"(v.y) : "l"(p));
	return v;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p)
{
	uint64_t v;
	asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream_u64x2(uint64_t *p, uint64_t a, uint64_t b)
{
	asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_stream_u64(uint64_t *p, uint64_t v)
{
#if defined(MSB64_PLAIN_STORES)
	*p = v;
#elif defined(MSB64_CS_STORES)
	__stcs(reinterpret_cast<unsigned long long *>(p), v);
#else
	asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
#endif
}

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
	asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		     :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\t"
			     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			     "selp.u32 %0, 1, 0, p;\n\t}"
			     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
	} while (!done);
}
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		     :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// Inclusive-to-exclusive block scan helper over one value per thread.
// scratch: THREADS/32 + 1 words of shared memory.  Returns the exclusive prefix;
// *total receives the block sum.  Contains two __syncthreads().
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *scratch,
							 uint32_t *total)
{
	constexpr int WARPS = THREADS / 32;
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += t;
	}
	if (lane == 31) scratch[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < WARPS ? scratch[lane] : 0;
		uint32_t winc = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc += t;
		}
		if (lane < WARPS) scratch[lane] = winc - w;
		if (lane == 31) scratch[WARPS] = winc;
	}
	__syncthreads();
	*total = scratch[WARPS];
	return inc - v + scratch[warp];
}

} // namespace msb64
