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
// The kernel is laid out around the unit's 16-BYTE ALIGNED WINDOW to cut the instruction count
// (it is issue- and latency-bound, not memory-bound):
//   * a unit [begin, begin + size) is seen through the window that starts at the even element
//     begin - (begin & 1); thread t owns the 16-byte chunks t, t + THREADS, ... of the window:
//     keys arrive by 16-byte loads (half the load instructions and address arithmetic);
//   * a pair's slot number is its WINDOW index, the packed words are stored at window
//     positions (bin bases start at begin & 1), so the write-out reads two neighbouring words
//     with one 16-byte shared load and writes keys and rids with 16-byte stores;
//   * the bin table is laid out plainly (bin d at word d): one address computation per
//     counter access instead of the swizzle the conflict-free scan needed (the scan now takes
//     a 2-way bank conflict on its two 16-byte accesses per thread -- far cheaper than ~10
//     extra instructions per pair);
//   * the bin table is cleared during the write-out of the previous unit.
// Units are at most UNIT_CAP = LOCAL_CAP - 2 pairs, so the window never exceeds LOCAL_CAP slots.
#pragma once
#include "msb64_local_sort.cuh"
#include "msb64_scatter.cuh"     // mbarrier / bulk-copy primitives

namespace msb64 {

constexpr int PACK_SLOT_BITS = 12;
static_assert(LOCAL_CAP <= (1u << PACK_SLOT_BITS), "slot number must fit");
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


constexpr int PACK_CHUNKS = LOCAL_CAP / 2 / LOCAL_THREADS;       // 16-byte chunks per thread
static_assert(PACK_CHUNKS * 2 * LOCAL_THREADS == LOCAL_CAP, "window = chunks x threads");
constexpr int PACK_PER = LOCAL_NBINS / LOCAL_OWNERS;             // bins per scanning thread
static_assert(PACK_PER == 8 || PACK_PER == 4 || PACK_PER == 16, "scan reads whole 16-byte chunks");
constexpr size_t PACKED_SMEM = size_t(LOCAL_CAP) * 8                       // packed words, by window position
				+ size_t(LOCAL_CAP) * 8                     // rids as they landed, by window position
				+ (size_t(LOCAL_NBINS) + 32) * 4
				+ LOCAL_LIST_MAX * 4 + LOCAL_BIG_MAX * 4 + 64 * 4
				+ 2 * (LOCAL_THREADS / 32) * 8
				+ 16;                                       // mbarrier

__global__ void __launch_bounds__(LOCAL_THREADS, LOCAL_MINB)
local_sort_packed_kernel(const Ctx c, const uint64_t base_key)
{
	constexpr int THREADS = LOCAL_THREADS, CH = PACK_CHUNKS, WARPS = THREADS / 32, PER = PACK_PER;
	// slots are window indices < LOCAL_CAP; the mask also keeps the index read from a window
	// position outside the unit (stale bits, never stored) inside the rid window
	static_assert((LOCAL_CAP & (LOCAL_CAP - 1)) == 0, "slot mask");
	constexpr uint32_t SLOT_MASK = LOCAL_CAP - 1;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw);            // [LOCAL_CAP] packed words
	uint64_t *rin = pk + LOCAL_CAP;                                   // [LOCAL_CAP] rids of the window
	uint32_t *bins = reinterpret_cast<uint32_t *>(rin + LOCAL_CAP);   // [LOCAL_NBINS + 32]
	uint32_t *list = bins + LOCAL_NBINS + 32;                         // short bins to order: base | size << 16
	uint32_t *big = list + LOCAL_LIST_MAX;                            // long bins
	uint32_t *scratch = big + LOCAL_BIG_MAX;                          // [64]
	uint64_t *wred = reinterpret_cast<uint64_t *>(scratch + 64);      // [2 * WARPS] OR, AND per warp
	uint64_t *bar = wred + 2 * WARPS;
	__shared__ uint32_t s_nbig;

	const uint32_t tid = threadIdx.x, lane = lane_id();
	const uint32_t nunits = min(c.ctl->nunits, c.max_units);
	if (blockIdx.x >= nunits) return;
	if (tid == 0) {
		mbar_init(bar, 1);
		s_nbig = 0;
	}
	{
		uint4 *b4 = reinterpret_cast<uint4 *>(bins);
		for (uint32_t i = tid; i < (LOCAL_NBINS + 32) / 4; i += THREADS) b4[i] = make_uint4(0u, 0u, 0u, 0u);
	}

	// the unit's keys, chunk jj of thread tid = window slots 2 * (jj * THREADS + tid) + {0, 1};
	// chunks past the window re-read its last chunk (no divergence; they are never counted)
	ulonglong2 kx[CH];
	auto load_keys = [&](const Unit &x) {
		const uint32_t a = x.begin & 1u, w0 = x.begin - a, nch = (a + x.size + 1u) >> 1;
		const uint64_t *src = (x.buf ? c.keys[1] : c.keys[0]) + w0;
#pragma unroll
		for (int jj = 0; jj < CH; ++jj)
			if (uint32_t(jj * THREADS) < nch) {
				const uint32_t q = min(uint32_t(jj * THREADS) + tid, nch - 1);
				if (w0 + 2 * q + 1 < c.end) kx[jj] = ld_stream_u64x2(src + 2 * q);
				else kx[jj].x = kx[jj].y = ld_stream_u64(src + 2 * q);      // the array's last element
			}
	};
	Unit next = c.units[blockIdx.x];
	load_keys(next);
	uint32_t parity = 0;

	for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
		const Unit un = next;
		const bool more = u + gridDim.x < nunits;
		if (more) next = c.units[u + gridDim.x];                 // descriptor now, keys when the registers are free
		const uint32_t begin = un.begin, size = un.size;
		const uint32_t a = begin & 1u, w0 = begin - a;
		const uint32_t nch = (a + size + 1u) >> 1;               // 16-byte chunks of the window
		// how much of the rids' window a bulk copy may fetch without leaving the array (block-uniform)
		uint32_t elems = 2 * nch;
		const bool tail = w0 + elems > c.end;
		if (tail) elems -= 2;
		const uint64_t origin = unit_origin_key(un.origin) + ((un.origin & UNIT_LEVEL0) ? base_key : 0ull);

		// 1. OR / AND of key - origin over the unit's keys (a chunk's slot outside the unit takes
		//    its neighbour's key: every chunk a thread holds has at least one slot inside)
		uint64_t vor = 0, vand = ~0ull;
#pragma unroll
		for (int jj = 0; jj < CH; ++jj)
			if (uint32_t(jj * THREADS) < nch) {
				const uint32_t q = min(uint32_t(jj * THREADS) + tid, nch - 1);
				const bool in0 = 2 * q - a < size, in1 = 2 * q + 1 - a < size;
				const uint64_t e0 = (in0 ? kx[jj].x : kx[jj].y) - origin, e1 = (in1 ? kx[jj].y : kx[jj].x) - origin;
				vor |= e0 | e1;
				vand &= e0 & e1;
			}
		or_and_warp_to_shared<WARPS>(vor, vand, wred);
		__syncthreads();
		// every thread is past the previous unit's write-out: its rids may be overwritten
		if (tid == 0) {
			const uint64_t *src_rids = (un.buf ? c.rids[1] : c.rids[0]) + w0;
			if (elems) {
				mbar_expect_tx(bar, elems * 8);
				bulk_copy_g2s(rin, src_rids, elems * 8, bar);
			}
			if (tail)                                        // the window's last chunk: only the slots inside the array
				for (uint32_t w = elems; w < a + size; ++w) rin[w] = src_rids[w];
		}
		or_and_from_shared<WARPS>(wred, &vor, &vand);
		const uint64_t diff = vor & ~vand;
		uint64_t *out_keys = c.keys[0] + w0, *out_rids = c.rids[0] + w0;
		// window chunk q home: slots outside the unit belong to the neighbours and stay untouched
		auto store_chunk = [&](uint32_t q, uint64_t k0, uint64_t k1, uint64_t r0, uint64_t r1) {
			const bool in0 = 2 * q - a < size, in1 = 2 * q + 1 - a < size;
			if (in0 && in1) {
				st_stream_u64x2(out_keys + 2 * q, k0, k1);
				st_stream_u64x2(out_rids + 2 * q, r0, r1);
			} else if (in0) {
				st_stream_u64(out_keys + 2 * q, k0);
				st_stream_u64(out_rids + 2 * q, r0);
			} else if (in1) {
				st_stream_u64(out_keys + 2 * q + 1, k1);
				st_stream_u64(out_rids + 2 * q + 1, r1);
			}
		};

		if (diff == 0) {
			// all keys equal: nothing to order, only bring the pairs home
			if (elems) {
				mbar_wait(bar, parity);
				parity ^= 1u;
			}
			__syncthreads();                                   // thread 0's plain stores into rin
			if (un.buf != 0) {
#pragma unroll
				for (int jj = 0; jj < CH; ++jj) {
					const uint32_t q = jj * THREADS + tid;
					if (q < nch) store_chunk(q, kx[jj].x, kx[jj].y, rin[2 * q], rin[2 * q + 1]);
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
		const bool resolved = (diff & ((1ull << shift) - 1)) == 0;

		// 2. arrival ranks (branch-free inside a chunk row; slots outside the unit -> dummy bins)
		uint32_t rk[CH];
		uint32_t *dummy = bins + LOCAL_NBINS + lane;
#pragma unroll
		for (int jj = 0; jj < CH; ++jj) {
			rk[jj] = 0;
			if (uint32_t(jj * THREADS) < nch) {
				const uint32_t q = jj * THREADS + tid;
				kx[jj].x = (kx[jj].x - origin) & mask;              // from here on: the key's varying bits
				kx[jj].y = (kx[jj].y - origin) & mask;
				uint32_t *s0 = 2 * q - a < size ? &bins[uint32_t(kx[jj].x >> shift)] : dummy;
				uint32_t *s1 = 2 * q + 1 - a < size ? &bins[uint32_t(kx[jj].y >> shift)] : dummy;
				const uint32_t r0 = atomicAdd(s0, 1u), r1 = atomicAdd(s1, 1u);
				rk[jj] = r0 | (r1 << 16);
			}
		}
		__syncthreads();
		// scan over the bins (thread t owns bins t * PER ...); bins with 2..LOCAL_SERIAL_MAX keys are
		// numbered into `list` by the same scan; positions are window positions: they start at a
		uint32_t nlist;
		{
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			const bool own = tid < LOCAL_OWNERS && tid * PER < nb;
			uint32_t cn[PER];
#pragma unroll
			for (int ch = 0; ch < PER / 4; ++ch) {
				uint4 v = make_uint4(0u, 0u, 0u, 0u);
				if (own) v = b4[tid * (PER / 4) + ch];
				cn[4 * ch] = v.x;
				cn[4 * ch + 1] = v.y;
				cn[4 * ch + 2] = v.z;
				cn[4 * ch + 3] = v.w;
			}
			uint32_t sum = 0;
#pragma unroll
			for (int q = 0; q < PER; ++q) {
				sum += cn[q];
				if (!resolved && cn[q] - 2u <= LOCAL_SERIAL_MAX - 2u) sum += 1u << 16;
			}
			uint32_t total;
			const uint32_t ex = block_exclusive_scan<THREADS>(sum, scratch, &total);
			nlist = total >> 16;
			if (own) {
				uint32_t bbase = (ex & 0xffffu) + a, at = ex >> 16;
#pragma unroll
				for (int q = 0; q < PER; ++q) {
					const uint32_t o = bbase | (cn[q] << 16);
					bbase += cn[q];
					if (!resolved && cn[q] >= 2u) {
						if (cn[q] <= LOCAL_SERIAL_MAX) list[at++] = o;
						else big[atomicAdd(&s_nbig, 1u)] = tid * PER + q;
					}
					cn[q] = o;
				}
#pragma unroll
				for (int ch = 0; ch < PER / 4; ++ch)
					b4[tid * (PER / 4) + ch] = make_uint4(cn[4 * ch], cn[4 * ch + 1], cn[4 * ch + 2], cn[4 * ch + 3]);
			}
		}
		__syncthreads();

		// 3a. every word to bin base + arrival rank
#pragma unroll
		for (int jj = 0; jj < CH; ++jj)
			if (uint32_t(jj * THREADS) < nch) {
				const uint32_t q = jj * THREADS + tid;
				if (2 * q - a < size)
					pk[(bins[uint32_t(kx[jj].x >> shift)] & 0xffffu) + (rk[jj] & 0xffffu)] = (kx[jj].x << PACK_SLOT_BITS) | (2 * q);
				if (2 * q + 1 - a < size)
					pk[(bins[uint32_t(kx[jj].y >> shift)] & 0xffffu) + (rk[jj] >> 16)] = (kx[jj].y << PACK_SLOT_BITS) | (2 * q + 1);
			}
		__syncthreads();
		const uint32_t nbig = s_nbig;
		if (nbig) {
			// long bins: all keys equal -> nothing to order; otherwise flag the bin (bit 31)
#pragma unroll
			for (int jj = 0; jj < CH; ++jj)
				if (uint32_t(jj * THREADS) < nch) {
					const uint32_t q = jj * THREADS + tid;
#pragma unroll
					for (int e = 0; e < 2; ++e) {
						const uint64_t v = e ? kx[jj].y : kx[jj].x;
						if (2 * q + e - a < size) {
							const uint32_t d = uint32_t(v >> shift);
							const uint32_t w = bins[d];
							if (((w >> 16) & 0x7fffu) > LOCAL_SERIAL_MAX && !(w >> 31) &&
							    (pk[w & 0xffffu] >> PACK_SLOT_BITS) != v)
								atomicOr(&bins[d], 1u << 31);
						}
					}
				}
			__syncthreads();
			if (tid < nbig) big[tid] = bins[big[tid]];
			if (tid == 0) s_nbig = 0;
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

		// 4. home, a window chunk at a time: key = word's key bits + base, rid = the rid that landed
		//    at the word's slot; the bin table is cleared for the next unit on the way
		if (elems) {
			mbar_wait(bar, parity);
			parity ^= 1u;
		}
		{
			const ulonglong2 *pk2 = reinterpret_cast<const ulonglong2 *>(pk);
#pragma unroll
			for (int jj = 0; jj < CH; ++jj)
				if (uint32_t(jj * THREADS) < nch) {
					const uint32_t q = jj * THREADS + tid;
					if (q < nch) {
						const ulonglong2 x = pk2[q];
						store_chunk(q, (x.x >> PACK_SLOT_BITS) + base, (x.y >> PACK_SLOT_BITS) + base,
							    rin[uint32_t(x.x) & SLOT_MASK], rin[uint32_t(x.y) & SLOT_MASK]);
					}
				}
			uint4 *b4 = reinterpret_cast<uint4 *>(bins);
			for (uint32_t i = tid; i < (nb + 32) / 4; i += THREADS) b4[i] = make_uint4(0u, 0u, 0u, 0u);
			if (nb < LOCAL_NBINS && tid < 32) bins[LOCAL_NBINS + tid] = 0;
		}
	}
}

} // namespace msb64
