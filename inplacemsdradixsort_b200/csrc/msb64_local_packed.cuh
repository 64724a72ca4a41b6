// msb64_local_packed.cuh -- the local sort's fast path: one 64-bit word per pair.
//
// After the MSD levels the keys of a unit agree on everything above bit shift_L + bits_L
// (msb64_plan.cuh), so key - origin needs at most shift_L + bits_L bits.  When that leaves
// room for the pair's 12-bit slot number (shift_L + bits_L <= 52: every unit of a sort of
// more than ~2^24 uniform keys), the kernel sorts PACKED words
//
//        (key - base) << 12 | slot                       slot = position in the unit as loaded
//
// instead of moving keys and rids side by side: ordering the packed words orders the keys
// (ties by slot), the rid never moves -- it lands in shared memory once, by a bulk
// asynchronous copy (cp.async.bulk + mbarrier) that runs in the background while the keys
// are being sorted, and is gathered through the slot number on the way out.  Against the
// general kernel (msb64_local_sort.cuh) that halves the scattered shared-memory stores, the
// traffic of the collision fix-up and the registers held per pair.
//
// Steps are those of the general kernel: OR/AND of key - origin gives the differing bits;
// counting sort on the top differing bits with 1-2 bins per key (shared atomic = arrival
// rank, block scan = bin bases, short colliding bins listed by the same scan); one thread
// per listed bin orders its words (network up to 4, insertion above); long bins: all equal
// -> nothing to do, otherwise block-wide bitonic network; write-out: key = word >> 12 + base,
// rid = rin[slot].
#pragma once
#include "msb64_local_sort.cuh"
#include "msb64_scatter.cuh"     // mbarrier / bulk-copy primitives

namespace msb64 {

constexpr int PACK_SLOT_BITS = 12;
static_assert(LOCAL_CAP <= (1u << PACK_SLOT_BITS), "slot number must fit");
constexpr size_t PACKED_SMEM = size_t(LOCAL_CAP) * 8                       // packed words
			       + (size_t(LOCAL_CAP) + 2) * 8                 // rids as they landed (16-byte aligned window)
			       + (size_t(LOCAL_NBINS) + 32) * 4
			       + LOCAL_LIST_MAX * 4 + LOCAL_BIG_MAX * 4 + 64 * 4
			       + 2 * (LOCAL_THREADS / 32) * 8
			       + 16;                                         // mbarrier

// can a unit with this origin word take the packed path?
__host__ __device__ inline bool unit_packable(uint32_t origin)
{
	const uint32_t bits = (origin >> 6) & 15u;
	return bits != 0 && (origin & 63u) + bits + PACK_SLOT_BITS <= 64;
}

__device__ __forceinline__ void block_bitonic1(uint64_t *k, const uint32_t n)
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
					}
				}
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(LOCAL_THREADS, LOCAL_MINB)
local_sort_packed_kernel(const Ctx c, const uint64_t base_key)
{
	constexpr int THREADS = LOCAL_THREADS, ITEMS = LOCAL_ITEMS, WARPS = THREADS / 32;
	constexpr int OWNERS = LOCAL_OWNERS;
	constexpr uint32_t SLOT_MASK = (1u << PACK_SLOT_BITS) - 1;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw);            // [LOCAL_CAP] packed words
	uint64_t *rin = pk + LOCAL_CAP;                                   // [LOCAL_CAP + 2] rids, as loaded
	uint32_t *bins = reinterpret_cast<uint32_t *>(rin + LOCAL_CAP + 2);// [LOCAL_NBINS + 32]
	uint32_t *list = bins + LOCAL_NBINS + 32;                         // short bins to order: base | size << 16
	uint32_t *big = list + LOCAL_LIST_MAX;                            // long bins
	uint32_t *scratch = big + LOCAL_BIG_MAX;                          // [64]
	uint64_t *wred = reinterpret_cast<uint64_t *>(scratch + 64);      // [2 * WARPS] OR, AND per warp
	uint64_t *bar = wred + 2 * WARPS;
	__shared__ uint32_t s_nbig;

	const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
	const uint32_t nunits = min(c.ctl->nunits, c.max_units);
	if (blockIdx.x >= nunits) return;
	if (tid == 0) mbar_init(bar, 1);

	// software-pipelined over units: the keys of the next unit are requested as soon as
	// the current unit's words sit in shared memory
	uint64_t k[ITEMS];
	auto load_keys = [&](const Unit &x) {
		const uint64_t *src_keys = (x.buf ? c.keys[1] : c.keys[0]) + x.begin;
		const uint32_t nrows = (x.size + THREADS - 1) / THREADS;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < nrows) k[j] = ld_stream_u64(src_keys + min(uint32_t(j * THREADS) + tid, x.size - 1));
	};
	Unit next = c.units[blockIdx.x];
	load_keys(next);
	uint32_t parity = 0;

	for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
		const Unit un = next;
		const bool more = u + gridDim.x < nunits;
		if (more) next = c.units[u + gridDim.x];                 // descriptor now, keys when the registers are free
		const uint32_t begin = un.begin, size = un.size;
		const uint32_t rows = (size + THREADS - 1) / THREADS;
		// the rids' 16-byte aligned window [begin - a, ...) and how much of it a bulk copy may
		// fetch without leaving the array (block-uniform)
		const uint32_t a = begin & 1u;
		uint32_t elems = (a + size + 1u) & ~1u;
		const bool tail = (begin - a) + elems > c.end;             // the window's last slot is past the array
		if (tail) elems -= 2;
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			for (uint32_t i = tid; i < (LOCAL_NBINS + 32) / 4; i += THREADS)
				b4[i] = make_uint4(0u, 0u, 0u, 0u);
		}
		if (tid == 0) s_nbig = 0;
		const uint64_t origin = unit_origin_key(un.origin) + ((un.origin & UNIT_LEVEL0) ? base_key : 0ull);
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
		// every thread is past the previous unit's write-out: its rids may be overwritten
		if (tid == 0) {
			const uint64_t *src_rids = (un.buf ? c.rids[1] : c.rids[0]);
			if (elems) {
				mbar_expect_tx(bar, elems * 8);
				bulk_copy_g2s(rin, src_rids + (begin - a), elems * 8, bar);
			}
			if (tail) {
				// the last one or two pairs of the array: plain loads
				for (uint32_t i = elems > a ? elems - a : 0; i < size; ++i) rin[a + i] = src_rids[begin + i];
			}
		}
		{
			const uint64_t o = lane < WARPS ? wred[lane] : 0ull;
			const uint64_t d = lane < WARPS ? wred[WARPS + lane] : ~0ull;
			const uint32_t olo = __reduce_or_sync(0xffffffffu, uint32_t(o));
			const uint32_t ohi = __reduce_or_sync(0xffffffffu, uint32_t(o >> 32));
			const uint32_t alo = __reduce_and_sync(0xffffffffu, uint32_t(d));
			const uint32_t ahi = __reduce_and_sync(0xffffffffu, uint32_t(d >> 32));
			vor = (uint64_t(ohi) << 32) | olo;
			vand = (uint64_t(ahi) << 32) | alo;
		}
		const uint64_t diff = vor & ~vand;

		if (diff == 0) {
			// all keys equal: nothing to order, only bring the pairs home
			if (elems) {
				mbar_wait(bar, parity);
				parity ^= 1u;
			}
			__syncthreads();                                   // thread 0's plain stores into rin
			if (un.buf != 0) {
#pragma unroll
				for (int j = 0; j < ITEMS; ++j) {
					const uint32_t i = j * THREADS + tid;
					if (i < size) {
						st_stream_u64(c.keys[0] + begin + i, k[j]);
						st_stream_u64(c.rids[0] + begin + i, rin[a + i]);
					}
				}
			}
			if (more) load_keys(next);
			continue;
		}

		const int top = 63 - __clzll(diff);                      // highest differing bit
		if (top + 1 + PACK_SLOT_BITS > 64 && tid == 0) atomicOr(&c.ctl->error, 16u);   // the plan kernel's promise
		const uint64_t mask = (2ull << top) - 1;
		const uint64_t base = (vand & ~mask) + origin;           // key = base + (key - origin) & mask
		int b = 32 - __clz(size - 1);                            // ceil(log2(size)), size >= 2 here
		b = min(max(b + LOCAL_EXTRA_BITS, 5), LOCAL_BITS);
		b = min(b, top + 1);
		const int shift = top + 1 - b;
		const uint32_t nb = 1u << b;
		constexpr int LPER = LOCAL_LPER, CH = LOCAL_CHUNKS;
#ifdef MSB64_NATURAL_BINS
#define MSB64_BIN_SLOT(d) (d)
#define MSB64_BIN_CHUNK(ch, t) ((t) * CH + (ch))
#else
#define MSB64_BIN_SLOT(d) ((((((d) >> 2) & (CH - 1)) * OWNERS + ((d) >> LPER)) << 2) | ((d) & 3u))
#define MSB64_BIN_CHUNK(ch, t) ((ch) * OWNERS + (t))
#endif
		const bool resolved = (diff & ((1ull << shift) - 1)) == 0;

		// 2. arrival ranks (branch-free inside a row; slots past the end -> dummy bins)
		uint32_t rk[ITEMS / 2];
#pragma unroll
		for (int j = 0; j < ITEMS / 2; ++j) rk[j] = 0;
		uint32_t *dummy = bins + LOCAL_NBINS + lane;
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) {
				const uint32_t i = j * THREADS + tid;
				k[j] = (k[j] - origin) & mask;                      // from here on: the key's varying bits
				const uint32_t d = uint32_t(k[j] >> shift);
				uint32_t *slot = i < size ? &bins[MSB64_BIN_SLOT(d)] : dummy;
				rk[j >> 1] |= atomicAdd(slot, 1u) << (16 * (j & 1));
			}
		__syncthreads();
		// scan over the bins; bins with 2..LOCAL_SERIAL_MAX keys are numbered into `list`
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
				uint32_t bbase = ex & 0xffffu, at = ex >> 16;
#pragma unroll
				for (int q = 0; q < 4 * CH; ++q) {
					const uint32_t o = bbase | (cn[q] << 16);
					bbase += cn[q];
					if (!resolved && cn[q] >= 2u) {
						if (cn[q] <= LOCAL_SERIAL_MAX) list[at++] = o;
						else big[atomicAdd(&s_nbig, 1u)] = uint32_t((MSB64_BIN_CHUNK(q >> 2, tid) << 2) | (q & 3));
					}
					cn[q] = o;
				}
#pragma unroll
				for (int ch = 0; ch < CH; ++ch)
					b4[MSB64_BIN_CHUNK(ch, tid)] = make_uint4(cn[4 * ch], cn[4 * ch + 1], cn[4 * ch + 2], cn[4 * ch + 3]);
			}
		}
		__syncthreads();

		// 3a. every word to bin base + arrival rank
#pragma unroll
		for (int j = 0; j < ITEMS; ++j)
			if (j < rows) {
				const uint32_t i = j * THREADS + tid;
				if (i < size) {
					const uint32_t d = uint32_t(k[j] >> shift);
					const uint32_t p = (bins[MSB64_BIN_SLOT(d)] & 0xffffu) + ((rk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
					pk[p] = (k[j] << PACK_SLOT_BITS) | i;
				}
			}
		__syncthreads();
		const uint32_t nbig = s_nbig;
		if (nbig) {
			// long bins: all keys equal -> nothing to order; otherwise flag the bin (bit 31)
#pragma unroll
			for (int j = 0; j < ITEMS; ++j)
				if (j < rows) {
					const uint32_t i = j * THREADS + tid;
					if (i < size) {
						const uint32_t d = uint32_t(k[j] >> shift);
						const uint32_t e = bins[MSB64_BIN_SLOT(d)];
						if (((e >> 16) & 0x7fffu) > LOCAL_SERIAL_MAX && !(e >> 31) &&
						    (pk[e & 0xffffu] >> PACK_SLOT_BITS) != k[j])
							atomicOr(&bins[MSB64_BIN_SLOT(d)], 1u << 31);
					}
				}
			__syncthreads();
			if (tid < nbig) big[tid] = bins[big[tid]];
		}
		// the registers are free: request the next unit's keys now
		if (more) load_keys(next);
		// 3b. one thread per short colliding bin: words only, the rids stay where they are
		for (uint32_t q = tid; q < nlist; q += THREADS) {
			const uint32_t e = list[q];
			uint64_t *bk = pk + (e & 0xffffu);
			const uint32_t cnt = e >> 16;
			if (cnt == 2) {
				const uint64_t a0 = bk[0], a1 = bk[1];
				if (a0 > a1) {
					bk[0] = a1;
					bk[1] = a0;
				}
				continue;
			}
			if (cnt <= 4) {
				uint64_t a0 = bk[0], a1 = bk[1], a2 = bk[2], a3 = cnt > 3 ? bk[3] : ~0ull;
#define MSB64_CE1(x, y) { const uint64_t lo_ = x < y ? x : y, hi_ = x < y ? y : x; x = lo_; y = hi_; }
				MSB64_CE1(a0, a1)
				MSB64_CE1(a2, a3)
				MSB64_CE1(a0, a2)
				MSB64_CE1(a1, a3)
				MSB64_CE1(a1, a2)
#undef MSB64_CE1
				bk[0] = a0;
				bk[1] = a1;
				bk[2] = a2;
				if (cnt > 3) bk[3] = a3;
				continue;
			}
			for (uint32_t i = 1; i < cnt; ++i) {
				const uint64_t key = bk[i];
				uint32_t at = i;
				while (at > 0 && bk[at - 1] > key) {
					bk[at] = bk[at - 1];
					--at;
				}
				bk[at] = key;
			}
		}
		__syncthreads();
		// 3c. long bins whose keys are not all equal (adversarial bit patterns): block-wide network
		for (uint32_t q = 0; q < nbig; ++q) {
			const uint32_t e = big[q];
			if (e >> 31) block_bitonic1(pk + (e & 0xffffu), (e >> 16) & 0x7fffu);
		}

		// 4. home: key = word's key bits + base, rid = the rid that was loaded at the word's slot
		if (elems) {
			mbar_wait(bar, parity);
			parity ^= 1u;
		}
		for (uint32_t i = tid; i < size; i += THREADS) {
			const uint64_t x = pk[i];
			st_stream_u64(c.keys[0] + begin + i, (x >> PACK_SLOT_BITS) + base);
			st_stream_u64(c.rids[0] + begin + i, rin[a + (uint32_t(x) & SLOT_MASK)]);
		}
#undef MSB64_BIN_SLOT
#undef MSB64_BIN_CHUNK
	}
}

} // namespace msb64
