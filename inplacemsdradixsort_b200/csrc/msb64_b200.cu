// msb64_b200.cu -- C ABI (include/msb64_b200.h) and host-side driver of the B200 MSD radix
// sort.  Compiled for sm_100a only; there is no CPU path in this file or behind it.
//
// Host flow of one device sort (msb64_b200_sort_device), all on one stream, no host
// synchronisation between kernels:
//
//   init                               control block, level-0 segment / tiles
//   for each digit (level) of the schedule:
//       histogram -> plan -> scatter   (kernels return at once when the level is empty; the
//                                       schedule fixes the digit WIDTH of a level, the digit's
//                                       position travels with every segment, msb64_plan.cuh)
//   local_sort (packed, general)       all small-bucket units of all levels
//   copy                               finished buckets that ended in the scratch buffer
//
// Reference map: sort() msb_64.c:2261-2430, local_radixsort msb_64.c:1007-1035,
// schedule_passes msb_64.c:1334-1400, check() msb_64.c:2432-2505.
#include "../../include/msb64_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "msb64_common.cuh"
#include "msb64_histogram.cuh"
#include "msb64_local_sort.cuh"
#include "msb64_local_packed.cuh"
#include "msb64_plan.cuh"
#include "msb64_scatter.cuh"
#include "msb64_route.cuh"

using namespace msb64;

// launch shape of the scatter kernel (tuned on B200)
#ifndef MSB64_SCATTER_THREADS
#define MSB64_SCATTER_THREADS 256
#endif
#ifndef MSB64_SCATTER_MINB
#define MSB64_SCATTER_MINB 3
#endif
constexpr int SCATTER_THREADS = MSB64_SCATTER_THREADS;
constexpr int SCATTER_MINB = MSB64_SCATTER_MINB;

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::mutex g_mutex;

int fail(int code, const char *fmt, const char *detail = "")
{
	snprintf(g_err, sizeof(g_err), fmt, detail);
	return code;
}

#define CUDA_TRY(expr)                                                                     \
	do {                                                                               \
		cudaError_t e_ = (expr);                                                   \
		if (e_ != cudaSuccess) {                                                   \
			snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, \
				 #expr, cudaGetErrorString(e_));                           \
			return e_ == cudaErrorMemoryAllocation ? MSB64_ERR_NOMEM           \
							       : MSB64_ERR_CUDA;           \
		}                                                                          \
	} while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ schedule
std::vector<int> g_schedule_override;
const bool g_no_fuse = getenv("MSB64_NO_FUSE") != nullptr;     // developer switch: separate histogram pass per level

// Digit widths, most significant first (the role of schedule_passes, msb_64.c:1334), for
// keys of which only the low `width` bits vary (64 when nothing is known about the keys).
std::vector<int> make_schedule(uint64_t n, int width = 64)
{
	if (width == 64 && !g_schedule_override.empty()) return g_schedule_override;
	// Uniform keys stop descending once the average bucket fits the local sort with
	// room to spare (2048 pairs): that takes `need` bits.  The scatter's HBM efficiency
	// falls with the run length TILE / 2^bits (tools/permcopy.cu: 6.3 TB/s at 256-byte
	// runs, 3.5 TB/s at 64-byte runs), so the `need` bits are spread over the fewest
	// passes of at most 8 bits, as evenly as possible, widest first.
	// log2(n) rounded to the nearest integer (a receive count just above a power of two
	// must not buy a whole extra bit: buckets of 2100 pairs are as good as 2048)
	int log_n = 0;
	while (log_n < 63 && (2ull << log_n) <= n) ++log_n;                       // floor(log2 n)
	if (log_n < 63 && double(n) > 1.41421356 * double(1ull << log_n)) ++log_n;
	int need = log_n > 11 ? log_n - 11 : 0;
	if (need > width) need = width;
	std::vector<int> s;
	int used = 0;
	if (need >= 4) {
		const int passes = (need + 7) / 8;
		for (int p = 0; p < passes; ++p) {
			int b = (need - used + (passes - p) - 1) / (passes - p);     // ceil of the even share
			b = b < 4 ? 4 : b;
			s.push_back(b);
			used += b;
		}
	}
	// the rest of the key (only skewed inputs get here): 7-bit digits, 4..7 bits at the end.
	// Dead digits cost next to nothing (the plan kernel moves a segment down to its highest
	// differing bit), so the digits that do get used should be narrow enough for full-speed
	// scatter passes.  The last digit may reach above `width` (those bits are equal in every key).
	int rest = width - used;
	while (rest > 0) {
		int b = rest <= 7 ? (rest < 4 ? 4 : rest) : (rest < 11 ? rest - rest / 2 : 7);
		s.push_back(b);
		rest -= b;
	}
	if (s.empty()) s.push_back(4);
	return s;
}

// Level-0 digit of a sort whose keys are known to lie in [lo, hi]: (key >> shift0) - origin0
// with the schedule made for the bits that actually vary.  Returns the schedule; the
// digits below level 0 are plain bit fields under shift0.
struct RangePlan {
	std::vector<int> sched;
	int shift0;
	uint64_t origin0;       // lo >> shift0
};

RangePlan plan_range(uint64_t n, uint64_t lo, uint64_t hi)
{
	RangePlan r;
	if (hi < lo) hi = lo;
	const uint64_t span = hi - lo;
	int width = 0;
	while (width < 64 && (span >> width)) ++width;
	if (width < 1) width = 1;
	for (;; ++width) {
		r.sched = make_schedule(n, width);
		const int bits0 = r.sched[0];
		r.shift0 = width > bits0 ? width - bits0 : 0;
		r.origin0 = lo >> r.shift0;
		// the digit of the largest key must fit: (hi >> shift0) - origin0 < 2^bits0
		if (width >= 64 || ((hi >> r.shift0) - r.origin0) < (1ull << bits0)) break;
	}
	return r;
}

// ------------------------------------------------------------------ device state
struct Device {
	bool ready = false;
	int sms = 0;
	int hist_blocks[MAX_BITS + 1] = {0};      // resident blocks per SM, by digit width
	int fused_blocks[MAX_BITS + 1] = {0};     // same for the fused (two-level) histogram
	int scatter_blocks[MAX_BITS + 1] = {0};
	int local_blocks = 0, packed_blocks = 0;
	// cached allocations (grow-only)
	void *ws = nullptr;
	size_t ws_bytes = 0;
	uint64_t *dkeys = nullptr, *drids = nullptr;
	size_t dcap = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4 * MAX_LEVELS + 16];
	bool events = false;
	// last sort
	uint64_t level_us[MAX_LEVELS][3] = {{0}};   // histogram, plan, scatter per level (timed sorts)
	int last_levels = 0;
	Control *last_ctl = nullptr;
	cudaStream_t last_stream = nullptr;
} g_dev;

template <int BITS>
int setup_bits()
{
	using H = HistCfg<BITS, 256>;
	using S = ScatterCfg<BITS, SCATTER_THREADS>;
	CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<BITS, 256, false>,
				      cudaFuncAttributeMaxDynamicSharedMemorySize, int(H::SMEM)));
	if (BITS < FUSE_MAX_BITS - 3) {
		CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<BITS, 256, true>,
					      cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(H::SMEM + (size_t(H::NB + 32) << (FUSE_MAX_BITS - BITS)) * 4)));
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
			&g_dev.fused_blocks[BITS], histogram_kernel<BITS, 256, true>, 256,
			H::SMEM + (size_t(H::NB + 32) << (FUSE_MAX_BITS - BITS)) * 4));
	}
	CUDA_TRY(cudaFuncSetAttribute(scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB>,
				      cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
		&g_dev.hist_blocks[BITS], histogram_kernel<BITS, 256, false>, 256, H::SMEM));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
		&g_dev.scatter_blocks[BITS], scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB>, SCATTER_THREADS, S::SMEM));
	if (g_dev.hist_blocks[BITS] < 1 || g_dev.scatter_blocks[BITS] < 1)
		return fail(MSB64_ERR_CUDA, "kernel does not fit on an SM%s");
	return MSB64_OK;
}

int device_init()
{
	if (g_dev.ready) return MSB64_OK;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(MSB64_ERR_CUDA, "no CUDA device: %s",
			    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
	if (prop.major < 10)
		return fail(MSB64_ERR_CUDA, "device %s is not sm_100 class", prop.name);
	g_dev.sms = prop.multiProcessorCount;
	int rc;
	if ((rc = setup_bits<4>()) || (rc = setup_bits<5>()) || (rc = setup_bits<6>()) ||
	    (rc = setup_bits<7>()) || (rc = setup_bits<8>()) || (rc = setup_bits<9>()) ||
	    (rc = setup_bits<10>()) || (rc = setup_bits<11>()))
		return rc;
	CUDA_TRY(cudaFuncSetAttribute(local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				      int(LOCAL_SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_dev.local_blocks, local_sort_kernel,
							       LOCAL_THREADS, LOCAL_SMEM));
	if (g_dev.local_blocks < 1) return fail(MSB64_ERR_CUDA, "local sort does not fit on an SM%s");
	CUDA_TRY(cudaFuncSetAttribute(local_sort_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				      int(PACKED_SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_dev.packed_blocks, local_sort_packed_kernel,
							       LOCAL_THREADS, PACKED_SMEM));
	if (g_dev.packed_blocks < 1) return fail(MSB64_ERR_CUDA, "packed local sort does not fit on an SM%s");
	CUDA_TRY(cudaStreamCreateWithFlags(&g_dev.stream, cudaStreamNonBlocking));
	g_dev.ready = true;
	return MSB64_OK;
}

int ensure_events()
{
	if (g_dev.events) return MSB64_OK;
	for (auto &ev : g_dev.ev) CUDA_TRY(cudaEventCreate(&ev));
	g_dev.events = true;
	return MSB64_OK;
}

// ------------------------------------------------------------------ workspace
struct Layout {
	size_t keys_b, rids_b, segs[2], tiles[2], hist[2], segbits[2], units, copies, ctl, fused, total;
	uint32_t max_segs, max_tiles, max_units, max_copies;
};

Layout make_layout(uint64_t n, const std::vector<int> &sched)
{
	Layout L;
	int maxbits = 4;
	for (int b : sched) maxbits = b > maxbits ? b : maxbits;
	const uint64_t levels = sched.size();
	L.max_segs = uint32_t(n / LOCAL_CAP + 2);
	L.max_tiles = uint32_t(n / TILE + 2 * uint64_t(L.max_segs) + 2);
	L.max_units = uint32_t(2 * (n / LOCAL_CAP) + 2 * levels * L.max_segs + 16);
	L.max_copies = uint32_t(n / COPY_TILE + L.max_segs + 2);
	size_t at = 0;
	auto take = [&](size_t bytes) { size_t o = at; at = align_up(at + bytes); return o; };
	L.keys_b = take(n * 8);
	L.rids_b = take(n * 8);
	for (int i = 0; i < 2; ++i) L.segs[i] = take(size_t(L.max_segs) * sizeof(Seg));
	for (int i = 0; i < 2; ++i) L.tiles[i] = take(size_t(L.max_tiles) * sizeof(Tile));
	for (int i = 0; i < 2; ++i) L.hist[i] = take((size_t(L.max_segs) << maxbits) * 4);
	for (int i = 0; i < 2; ++i) L.segbits[i] = take(size_t(L.max_segs) * sizeof(SegBits));
	L.units = take(size_t(L.max_units) * sizeof(Unit));
	L.copies = take(size_t(L.max_copies) * sizeof(CopyTile));
	L.ctl = take(sizeof(Control));
	L.fused = take((size_t(1) << FUSE_MAX_BITS) * 4);
	L.total = at;
	return L;
}

// ------------------------------------------------------------------ launches
template <int BITS>
void launch_level(const Ctx &c, int level, int shift0, uint32_t origin, int next_bits, cudaStream_t st,
		  cudaEvent_t *ev)
{
	using H = HistCfg<BITS, 256>;
	using S = ScatterCfg<BITS, SCATTER_THREADS>;
	if (ev) cudaEventRecord(ev[0], st);
	// level 0 is one segment: its histogram pass also counts the level-1 digits per bin
	// (32 KiB of shared counters), and level 1 needs no histogram pass
	const bool fuse = level == 0 && next_bits > 0 && BITS + next_bits <= FUSE_MAX_BITS &&
			  BITS < FUSE_MAX_BITS - 3 && g_dev.fused_blocks[BITS] > 0 && shift0 >= next_bits && !g_no_fuse;
	if (fuse) {
		const size_t smem = H::SMEM + (size_t(H::NB + 32) << next_bits) * 4;
		histogram_kernel<BITS, 256, true><<<g_dev.sms * g_dev.fused_blocks[BITS], 256, smem, st>>>(
			c, level, origin, next_bits);
	} else {
		histogram_kernel<BITS, 256, false><<<g_dev.sms * g_dev.hist_blocks[BITS], 256, H::SMEM, st>>>(
			c, level, origin, 0);
	}
	if (ev) cudaEventRecord(ev[1], st);
	plan_kernel<<<g_dev.sms * 4, PLAN_THREADS, 0, st>>>(c, level, BITS, next_bits, fuse);
	if (ev) cudaEventRecord(ev[2], st);
	scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB><<<g_dev.sms * g_dev.scatter_blocks[BITS], SCATTER_THREADS, S::SMEM, st>>>(c, level, origin);
	if (ev) cudaEventRecord(ev[3], st);
	g_launches += 3;
}

int sort_device_locked(uint64_t *d_keys, uint64_t *d_rids, uint64_t n, void *workspace,
		       size_t workspace_bytes, cudaStream_t st, uint64_t *phase_us,
		       uint64_t key_lo = 0, uint64_t key_hi = ~0ull)
{
	int rc = device_init();
	if (rc) return rc;
	if (n > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	if (phase_us) memset(phase_us, 0, MSB64_PHASE_COUNT * sizeof(uint64_t));
	if (n < 2) return MSB64_OK;
	if (!d_keys || !d_rids || (uintptr_t(d_keys) & 15) || (uintptr_t(d_rids) & 15))
		return fail(MSB64_ERR_ARG, "device arrays must be non-NULL and 16-byte aligned%s");

	const RangePlan rp = plan_range(n, key_lo, key_hi);
	const std::vector<int> &sched = rp.sched;
	const Layout L = make_layout(n, sched);
	if (!workspace) {
		if (g_dev.ws_bytes < L.total) {
			if (g_dev.ws) cudaFree(g_dev.ws);
			g_dev.ws = nullptr;
			g_dev.ws_bytes = 0;
			CUDA_TRY(cudaMalloc(&g_dev.ws, L.total));
			g_dev.ws_bytes = L.total;
		}
		workspace = g_dev.ws;
	} else if (workspace_bytes < L.total || (uintptr_t(workspace) & 255)) {
		return fail(MSB64_ERR_NOMEM, "workspace too small or not 256-byte aligned%s");
	}
	char *w = static_cast<char *>(workspace);
	Ctx c;
	c.keys[0] = d_keys;
	c.rids[0] = d_rids;
	c.keys[1] = reinterpret_cast<uint64_t *>(w + L.keys_b);
	c.rids[1] = reinterpret_cast<uint64_t *>(w + L.rids_b);
	for (int i = 0; i < 2; ++i) {
		c.segs[i] = reinterpret_cast<Seg *>(w + L.segs[i]);
		c.tiles[i] = reinterpret_cast<Tile *>(w + L.tiles[i]);
		c.hist[i] = reinterpret_cast<uint32_t *>(w + L.hist[i]);
		c.segbits[i] = reinterpret_cast<SegBits *>(w + L.segbits[i]);
	}
	c.units = reinterpret_cast<Unit *>(w + L.units);
	c.copies = reinterpret_cast<CopyTile *>(w + L.copies);
	c.ctl = reinterpret_cast<Control *>(w + L.ctl);
	c.fused = reinterpret_cast<uint32_t *>(w + L.fused);
	c.n = uint32_t(n);
	c.max_segs = L.max_segs;
	c.max_tiles = L.max_tiles;
	c.max_units = L.max_units;
	c.max_copies = L.max_copies;

	cudaEvent_t *ev = nullptr;
	if (phase_us) {
		if ((rc = ensure_events())) return rc;
		ev = g_dev.ev;
	}
	const int levels = int(sched.size());
	if (ev) cudaEventRecord(ev[0], st);
	init_kernel<<<g_dev.sms, 256, 0, st>>>(c, sched[0], rp.shift0);
	g_launches += 1;
	if (n > LOCAL_CAP) {
		// the position of a segment's digit travels with the segment (msb64_plan.cuh); the
		// host only says where level 0 starts
		for (int l = 0; l < levels; ++l) {
			const int bits = sched[l];
			const int shift = rp.shift0;
			const uint32_t origin = l == 0 ? uint32_t(rp.origin0) : 0u;
			const int next_bits = l + 1 < levels ? sched[l + 1] : 0;
			cudaEvent_t *lev = ev ? ev + 1 + 4 * l : nullptr;
			switch (bits) {
			case 4: launch_level<4>(c, l, shift, origin, next_bits, st, lev); break;
			case 5: launch_level<5>(c, l, shift, origin, next_bits, st, lev); break;
			case 6: launch_level<6>(c, l, shift, origin, next_bits, st, lev); break;
			case 7: launch_level<7>(c, l, shift, origin, next_bits, st, lev); break;
			case 8: launch_level<8>(c, l, shift, origin, next_bits, st, lev); break;
			case 9: launch_level<9>(c, l, shift, origin, next_bits, st, lev); break;
			case 10: launch_level<10>(c, l, shift, origin, next_bits, st, lev); break;
			case 11: launch_level<11>(c, l, shift, origin, next_bits, st, lev); break;
			default: return fail(MSB64_ERR_ARG, "digit width outside 4..11%s");
			}
		}
	}
	cudaEvent_t *tail = ev ? ev + 1 + 4 * levels : nullptr;
	if (tail) cudaEventRecord(tail[0], st);
	// units whose keys leave room for a slot number in one word take the packed path, the rest
	// (small arrays, very deep levels never) the general one; an empty list costs a launch
	local_sort_packed_kernel<<<g_dev.sms * g_dev.packed_blocks, LOCAL_THREADS, PACKED_SMEM, st>>>(
		c, rp.origin0 << rp.shift0);
	local_sort_kernel<<<g_dev.sms * g_dev.local_blocks, LOCAL_THREADS, LOCAL_SMEM, st>>>(
		c, rp.origin0 << rp.shift0);
	g_launches += 1;
	if (tail) cudaEventRecord(tail[1], st);
	copy_kernel<<<g_dev.sms * 8, 256, 0, st>>>(c);
	if (tail) cudaEventRecord(tail[2], st);
	g_launches += 2;
	CUDA_TRY(cudaGetLastError());
	g_dev.last_ctl = c.ctl;
	g_dev.last_stream = st;

	if (phase_us) {
		CUDA_TRY(cudaStreamSynchronize(st));
		auto us = [&](cudaEvent_t a, cudaEvent_t b) {
			float ms = 0;
			cudaEventElapsedTime(&ms, a, b);
			return uint64_t(ms * 1000.0f + 0.5f);
		};
		phase_us[MSB64_PHASE_PLAN] += us(ev[0], ev[1]);
		if (n > LOCAL_CAP)
			for (int l = 0; l < levels; ++l) {
				cudaEvent_t *lev = ev + 1 + 4 * l;
				g_dev.level_us[l][0] = us(lev[0], lev[1]);
				g_dev.level_us[l][1] = us(lev[1], lev[2]);
				g_dev.level_us[l][2] = us(lev[2], lev[3]);
				phase_us[MSB64_PHASE_HISTOGRAM] += g_dev.level_us[l][0];
				phase_us[MSB64_PHASE_PLAN] += g_dev.level_us[l][1];
				phase_us[MSB64_PHASE_SCATTER] += g_dev.level_us[l][2];
			}
		g_dev.last_levels = n > LOCAL_CAP ? levels : 0;
		phase_us[MSB64_PHASE_LOCAL] = us(tail[0], tail[1]);
		phase_us[MSB64_PHASE_COPY] = us(tail[1], tail[2]);
		Control h;
		CUDA_TRY(cudaMemcpy(&h, c.ctl, sizeof(h), cudaMemcpyDeviceToHost));
		if (h.error) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
	}
	return MSB64_OK;
}

// ------------------------------------------------------------------ fill / check kernels
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

__global__ void fill_kernel(uint64_t *keys, uint64_t *rids, uint64_t n, int kind, uint64_t seed,
			    uint64_t param)
{
	for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
	     i += uint64_t(gridDim.x) * blockDim.x) {
		const uint64_t x = mix64(seed * 0x9e3779b97f4a7c15ull + i + 1);
		uint64_t k;
		switch (kind) {
		case 1: k = x & param; break;
		case 2: k = mix64((x % (param ? param : 1)) + 0x51ed27ull); break;
		case 3: k = i * (param ? param : 1); break;
		case 4: k = (n - 1 - i) * (param ? param : 1); break;
		default: k = x;
		}
		keys[i] = k;
		if (rids) rids[i] = i;
	}
}

// descents, wrapping key sum, order-independent (key, rid) digest -- the device form of
// check() (msb_64.c:2432-2505); the digest formula is the oracle's orc_pair_digest.
__global__ void check_kernel(const uint64_t *keys, const uint64_t *rids, uint64_t n,
			     unsigned long long *out)
{
	unsigned long long bad = 0, sum = 0, dig = 0;
	for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
	     i += uint64_t(gridDim.x) * blockDim.x) {
		const uint64_t k = keys[i];
		if (i + 1 < n && k > keys[i + 1]) bad++;
		sum += k;
		if (rids) dig += mix64(k + 0x9e3779b97f4a7c15ull * mix64(rids[i] + 1));
	}
	for (int d = 16; d; d >>= 1) {
		bad += __shfl_xor_sync(0xffffffffu, bad, d);
		sum += __shfl_xor_sync(0xffffffffu, sum, d);
		dig += __shfl_xor_sync(0xffffffffu, dig, d);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&out[0], bad);
		atomicAdd(&out[1], sum);
		atomicAdd(&out[2], dig);
	}
}

// node boundaries of sort(): first index whose key exceeds the key at each quantile
__global__ void boundary_kernel(const uint64_t *keys, uint64_t n, int numa, uint64_t *bounds)
{
	const int node = threadIdx.x;
	if (node >= numa) return;
	if (node == numa - 1) {
		bounds[node] = n;
		return;
	}
	const uint64_t q = (n / numa) * (node + 1);
	if (q == 0) {
		bounds[node] = 0;
		return;
	}
	const uint64_t delim = keys[q - 1];
	uint64_t lo = q, hi = n;              // first index in [q, n) with key > delim
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (keys[mid] > delim) hi = mid;
		else lo = mid + 1;
	}
	bounds[node] = lo;
}

int ensure_device_arrays(uint64_t n)
{
	if (g_dev.dcap >= n) return MSB64_OK;
	if (g_dev.dkeys) cudaFree(g_dev.dkeys);
	if (g_dev.drids) cudaFree(g_dev.drids);
	g_dev.dkeys = g_dev.drids = nullptr;
	g_dev.dcap = 0;
	CUDA_TRY(cudaMalloc(&g_dev.dkeys, align_up(n * 8)));
	CUDA_TRY(cudaMalloc(&g_dev.drids, align_up(n * 8)));
	g_dev.dcap = n;
	return MSB64_OK;
}

const char *kPhaseNames[] = {
	"Copy to device time:      ", "Histogram time:           ", "Plan time:                ",
	"Scatter time:             ", "Local sort time:          ", "Copy home time:           ",
	"Copy to host time:        ",
};

int sort_host_locked(uint64_t **keys, uint64_t **rids, uint64_t *size, int numa, double fudge,
		     char **description, uint64_t *times)
{
	int rc = device_init();
	if (rc) return rc;
	if (!keys || !rids || !size || numa < 1 || numa > 64 || !(fudge >= 1.0))
		return fail(MSB64_ERR_ARG, "bad keys/rids/size/numa/fudge%s");
	uint64_t total = 0;
	for (int n = 0; n < numa; ++n) {
		if (size[n] && (!keys[n] || !rids[n])) return fail(MSB64_ERR_ARG, "NULL array%s");
		if ((uintptr_t(keys[n]) & 15) || (uintptr_t(rids[n]) & 15))
			return fail(MSB64_ERR_ARG, "arrays must be 16-byte aligned (msb_64.c:2272)%s");
		total += size[n];
	}
	if (total > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	if (description) description[0] = nullptr;
	if (total == 0) return MSB64_OK;
	if ((rc = ensure_device_arrays(total))) return rc;
	if ((rc = ensure_events())) return rc;
	cudaStream_t st = g_dev.stream;
	cudaEvent_t *ev = g_dev.ev + (4 * MAX_LEVELS + 8);     // spare events past the per-level ones

	CUDA_TRY(cudaEventRecord(ev[0], st));
	uint64_t at = 0;
	for (int n = 0; n < numa; ++n) {
		if (!size[n]) continue;
		CUDA_TRY(cudaMemcpyAsync(g_dev.dkeys + at, keys[n], size[n] * 8, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(g_dev.drids + at, rids[n], size[n] * 8, cudaMemcpyHostToDevice, st));
		at += size[n];
	}
	CUDA_TRY(cudaEventRecord(ev[1], st));
	uint64_t phase[MSB64_PHASE_COUNT];
	rc = sort_device_locked(g_dev.dkeys, g_dev.drids, total, nullptr, 0, st, times ? phase : nullptr);
	if (rc) return rc;

	// node boundaries: exact quantiles, equal keys never split (msb_64.c:1596-1606)
	std::vector<uint64_t> bounds(numa, total);
	if (numa > 1) {
		uint64_t *d_bounds = reinterpret_cast<uint64_t *>(g_dev.ws);   // B keys are dead now
		boundary_kernel<<<1, 64, 0, st>>>(g_dev.dkeys, total, numa, d_bounds);
		g_launches += 1;
		CUDA_TRY(cudaMemcpyAsync(bounds.data(), d_bounds, numa * 8, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		uint64_t prev = 0;
		for (int n = 0; n < numa; ++n) {
			const uint64_t cap = uint64_t(double(size[n]) * fudge);     // msb_64.c:1574
			if (bounds[n] - prev > cap)
				return fail(MSB64_ERR_CAPACITY, "a node would exceed size[n] * fudge%s");
			prev = bounds[n];
		}
	}
	CUDA_TRY(cudaEventRecord(ev[2], st));
	uint64_t prev = 0;
	for (int n = 0; n < numa; ++n) {
		const uint64_t cnt = bounds[n] - prev;
		if (cnt) {
			CUDA_TRY(cudaMemcpyAsync(keys[n], g_dev.dkeys + prev, cnt * 8, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaMemcpyAsync(rids[n], g_dev.drids + prev, cnt * 8, cudaMemcpyDeviceToHost, st));
		}
		size[n] = cnt;
		prev = bounds[n];
	}
	CUDA_TRY(cudaEventRecord(ev[3], st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (total >= 2) {
		Control h;
		CUDA_TRY(cudaMemcpy(&h, g_dev.last_ctl, sizeof(h), cudaMemcpyDeviceToHost));
		if (h.error) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
	}
	if (times && description) {
		float h2d = 0, d2h = 0;
		cudaEventElapsedTime(&h2d, ev[0], ev[1]);
		cudaEventElapsedTime(&d2h, ev[2], ev[3]);
		times[0] = uint64_t(h2d * 1000);
		for (int p = 0; p < MSB64_PHASE_COUNT; ++p) times[1 + p] = phase[p];
		times[1 + MSB64_PHASE_COUNT] = uint64_t(d2h * 1000);
		for (int p = 0; p < 2 + MSB64_PHASE_COUNT; ++p) description[p] = const_cast<char *>(kPhaseNames[p]);
		description[2 + MSB64_PHASE_COUNT] = nullptr;
	}
	return MSB64_OK;
}

} // namespace

// =================================================================== C ABI
extern "C" {

int msb64_b200_sort(uint64_t **keys, uint64_t **rids, uint64_t *size, int threads, int numa,
		    double fudge, char **description, uint64_t *times)
{
	(void) threads;
	std::lock_guard<std::mutex> lock(g_mutex);
	return sort_host_locked(keys, rids, size, numa, fudge, description, times);
}

void sort(uint64_t **keys, uint64_t **rids, uint64_t *size, int threads, int numa, double fudge,
	  char **description, uint64_t *times)
{
	const int rc = msb64_b200_sort(keys, rids, size, threads, numa, fudge, description, times);
	if (rc != MSB64_OK) {
		fprintf(stderr, "msb64_b200 sort(): error %d: %s\n", rc, g_err);
		abort();
	}
}

void *mamalloc(size_t size)
{
	void *ptr = nullptr;
	return posix_memalign(&ptr, 64, size) ? nullptr : ptr;
}

int msb64_b200_sort_host(uint64_t *keys, uint64_t *rids, uint64_t n)
{
	uint64_t *k[1] = {keys}, *r[1] = {rids}, s[1] = {n};
	return msb64_b200_sort(k, r, s, 64, 1, 1.0, nullptr, nullptr);
}

size_t msb64_b200_workspace_bytes(uint64_t n)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	return make_layout(n, make_schedule(n)).total;
}

int msb64_b200_sort_device(uint64_t *d_keys, uint64_t *d_rids, uint64_t n, void *workspace,
			   size_t workspace_bytes, void *stream, uint64_t *phase_us)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	return sort_device_locked(d_keys, d_rids, n, workspace, workspace_bytes,
				  static_cast<cudaStream_t>(stream), phase_us);
}

int msb64_b200_sort_device_range(uint64_t *d_keys, uint64_t *d_rids, uint64_t n, void *workspace,
				 size_t workspace_bytes, void *stream, uint64_t *phase_us,
				 uint64_t key_lo, uint64_t key_hi)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (key_hi < key_lo) return fail(MSB64_ERR_ARG, "key_hi < key_lo%s");
	return sort_device_locked(d_keys, d_rids, n, workspace, workspace_bytes,
				  static_cast<cudaStream_t>(stream), phase_us, key_lo, key_hi);
}

int msb64_b200_get_schedule(uint64_t n, int *bits)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	const std::vector<int> s = make_schedule(n);
	for (size_t i = 0; i < s.size(); ++i) bits[i] = s[i];
	return int(s.size());
}

int msb64_b200_get_range_schedule(uint64_t n, uint64_t key_lo, uint64_t key_hi, int *bits, int *shift0,
				  uint64_t *origin0)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (key_hi < key_lo || !bits) return 0;
	const RangePlan r = plan_range(n, key_lo, key_hi);
	for (size_t i = 0; i < r.sched.size(); ++i) bits[i] = r.sched[i];
	if (shift0) *shift0 = r.shift0;
	if (origin0) *origin0 = r.origin0;
	return int(r.sched.size());
}

int msb64_b200_set_schedule(const int *bits, int count)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (count == 0) {
		g_schedule_override.clear();
		return MSB64_OK;
	}
	if (count < 0 || count > MAX_LEVELS || !bits) return fail(MSB64_ERR_ARG, "bad schedule%s");
	int sum = 0;
	for (int i = 0; i < count; ++i) {
		if (bits[i] < 4 || bits[i] > MAX_BITS) return fail(MSB64_ERR_ARG, "digit width outside 4..11%s");
		sum += bits[i];
	}
	if (sum != 64) return fail(MSB64_ERR_ARG, "digit widths must sum to 64%s");
	g_schedule_override.assign(bits, bits + count);
	return MSB64_OK;
}

int msb64_b200_device_count(void)
{
	int count = 0;
	return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}

const char *msb64_b200_last_error(void) { return g_err; }

uint64_t msb64_b200_launch_count(void) { return g_launches.load(); }

int msb64_b200_last_stats(uint64_t *out, int cap)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (!g_dev.last_ctl) return 0;
	if (cudaStreamSynchronize(g_dev.last_stream) != cudaSuccess) return 0;
	Control h;
	if (cudaMemcpy(&h, g_dev.last_ctl, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
	int k = 0;
	for (int l = 0; l < MAX_LEVELS && k < cap; ++l) out[k++] = h.nsegs[l];
	for (int l = 0; l < MAX_LEVELS && k < cap; ++l) out[k++] = h.ntiles[l];
	if (k < cap) out[k++] = h.nunits + h.nslow;
	if (k < cap) out[k++] = h.ncopies;
	if (k < cap) out[k++] = h.error;
	if (k < cap) out[k++] = h.degenerate;
	if (k < cap) out[k++] = h.local_pairs;
	for (int l = 0; l < MAX_LEVELS && k < cap; ++l) out[k++] = h.moved[l];
	if (k < cap) out[k++] = h.hist_keys;
	return k;
}

int msb64_b200_last_level_times(uint64_t *out, int cap)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	int k = 0;
	for (int l = 0; l < g_dev.last_levels; ++l)
		for (int j = 0; j < 3 && k < cap; ++j) out[k++] = g_dev.level_us[l][j];
	return k;
}

int msb64_b200_digit_histogram(const uint64_t *d_keys, uint64_t n, int shift, int bits, uint64_t origin,
				uint64_t *d_hist, uint64_t *d_minmax, void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	int rc = device_init();
	if (rc) return rc;
	if (bits < 1 || bits > ROUTE_MAX_BITS || shift < 0 || shift >= 64 || !d_hist)
		return fail(MSB64_ERR_ARG, "digit_histogram: bad shift/bits%s");
	if (n > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(uint64_t) << bits, st));
	if (d_minmax) {
		CUDA_TRY(cudaMemsetAsync(d_minmax, 0xff, sizeof(uint64_t), st));
		CUDA_TRY(cudaMemsetAsync(d_minmax + 1, 0, sizeof(uint64_t), st));
	}
	if (n) {
		digit_histogram_kernel<<<g_dev.sms * 8, 256, sizeof(uint32_t) << bits, st>>>(
			d_keys, n, shift, bits, uint32_t(origin), reinterpret_cast<unsigned long long *>(d_hist),
			reinterpret_cast<unsigned long long *>(d_minmax));
		g_launches += 1;
	}
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

static int route_common(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, int shift, int bits,
			uint64_t origin, const uint8_t *d_bin_to_dest, int ndest, uint32_t *d_cursors, const RouteDst &dst,
			void *stream)
{
	int rc = device_init();
	if (rc) return rc;
	if (bits < 1 || bits > ROUTE_MAX_BITS || shift < 0 || shift >= 64 || ndest < 1 ||
	    ndest > ROUTE_MAX_DEST || !d_bin_to_dest || !d_cursors)
		return fail(MSB64_ERR_ARG, "route: bad shift/bits/ndest%s");
	if (n > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	if (!n) return MSB64_OK;
	static bool configured = false;
	if (!configured) {
		CUDA_TRY(cudaFuncSetAttribute(route_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(RouteCfg<16>::SMEM)));
		CUDA_TRY(cudaFuncSetAttribute(route_kernel<ROUTE_MAX_DEST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(RouteCfg<ROUTE_MAX_DEST>::SMEM)));
		configured = true;
	}
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (ndest <= 16)
		route_kernel<16><<<g_dev.sms * 3, ROUTE_THREADS, RouteCfg<16>::SMEM, st>>>(
			d_keys, d_rids, uint32_t(n), shift, bits, uint32_t(origin), d_bin_to_dest, ndest, d_cursors, dst);
	else
		route_kernel<ROUTE_MAX_DEST><<<g_dev.sms * 2, ROUTE_THREADS, RouteCfg<ROUTE_MAX_DEST>::SMEM, st>>>(
			d_keys, d_rids, uint32_t(n), shift, bits, uint32_t(origin), d_bin_to_dest, ndest, d_cursors, dst);
	g_launches += 1;
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

int msb64_b200_route(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, int shift, int bits,
		     uint64_t origin, const uint8_t *d_bin_to_dest, int ndest, uint32_t *d_cursors,
		     uint64_t *d_out_keys, uint64_t *d_out_rids, void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	RouteDst dst;
	for (int i = 0; i < ROUTE_MAX_DEST; ++i) {
		dst.keys[i] = d_out_keys;
		dst.rids[i] = d_out_rids;
	}
	return route_common(d_keys, d_rids, n, shift, bits, origin, d_bin_to_dest, ndest, d_cursors, dst, stream);
}

int msb64_b200_route_peer(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, int shift, int bits,
			  uint64_t origin, const uint8_t *d_bin_to_dest, int ndest, uint32_t *d_cursors,
			  uint64_t *const *out_keys, uint64_t *const *out_rids, void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (!out_keys || !out_rids || ndest < 1 || ndest > ROUTE_MAX_DEST)
		return fail(MSB64_ERR_ARG, "route_peer: bad destination table%s");
	RouteDst dst;
	for (int i = 0; i < ROUTE_MAX_DEST; ++i) {
		dst.keys[i] = out_keys[i < ndest ? i : 0];
		dst.rids[i] = out_rids[i < ndest ? i : 0];
	}
	return route_common(d_keys, d_rids, n, shift, bits, origin, d_bin_to_dest, ndest, d_cursors, dst, stream);
}

int msb64_b200_ipc_export(void *d_ptr, void *handle64)
{
	static_assert(sizeof(cudaIpcMemHandle_t) == MSB64_IPC_HANDLE_BYTES, "handle size");
	if (!d_ptr || !handle64) return fail(MSB64_ERR_ARG, "ipc_export: NULL%s");
	cudaIpcMemHandle_t h;
	CUDA_TRY(cudaIpcGetMemHandle(&h, d_ptr));
	memcpy(handle64, &h, sizeof(h));
	return MSB64_OK;
}

void *msb64_b200_ipc_open(const void *handle64)
{
	if (!handle64) return nullptr;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	void *p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) {
		snprintf(g_err, sizeof(g_err), "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

int msb64_b200_ipc_close(void *mapped)
{
	if (mapped) CUDA_TRY(cudaIpcCloseMemHandle(mapped));
	return MSB64_OK;
}

void *msb64_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void msb64_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

void *msb64_b200_device_alloc(size_t bytes)
{
	void *p = nullptr;
	return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
}
void msb64_b200_device_free(void *p) { if (p) cudaFree(p); }

int msb64_b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
	return MSB64_OK;
}
int msb64_b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
	return MSB64_OK;
}
int msb64_b200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
	CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
	return MSB64_OK;
}
int msb64_b200_stream_sync(void *stream)
{
	CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
	return MSB64_OK;
}

int msb64_b200_fill(uint64_t *d_keys, uint64_t *d_rids, uint64_t n, int kind, uint64_t seed,
		    uint64_t param, void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	int rc = device_init();
	if (rc) return rc;
	if (!n) return MSB64_OK;
	fill_kernel<<<g_dev.sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_keys, d_rids, n, kind,
										seed, param);
	g_launches += 1;
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

int msb64_b200_check(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, uint64_t *out,
		     void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	int rc = device_init();
	if (rc) return rc;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	unsigned long long *d_out = nullptr;
	CUDA_TRY(cudaMalloc(&d_out, 3 * sizeof(unsigned long long)));
	CUDA_TRY(cudaMemsetAsync(d_out, 0, 3 * sizeof(unsigned long long), st));
	if (n) {
		check_kernel<<<g_dev.sms * 8, 256, 0, st>>>(d_keys, d_rids, n, d_out);
		g_launches += 1;
	}
	cudaError_t e = cudaMemcpyAsync(out, d_out, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(d_out);
	CUDA_TRY(e);
	return MSB64_OK;
}

} // extern "C"
