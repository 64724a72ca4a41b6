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

#include <algorithm>
#include <atomic>
#include <chrono>
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
#include "msb64_tail.cuh"

using namespace msb64;

// launch shape of the scatter kernel (tuned on B200)
// 512 threads x 2 blocks per SM: 17.4 ms for the three passes of a 2^30 sort against 18.4 ms with
// 256 x 3 (more warps to cover one another's phases, 64 registers each; 512 x 3 spills: 20.9 ms)
#ifndef MSB64_SCATTER_THREADS
#define MSB64_SCATTER_THREADS 512
#endif
#ifndef MSB64_SCATTER_MINB
#define MSB64_SCATTER_MINB 2
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
const int g_tail_from = getenv("MSB64_TAIL_FROM") ? atoi(getenv("MSB64_TAIL_FROM")) : 0;   // developer switch: levels from here on run in the tail kernel (they must be 7 bits wide)
const bool g_debug_sync = getenv("MSB64_DEBUG_SYNC") != nullptr;   // developer switch: synchronise behind every kernel, name the one that faults

void debug_sync(cudaStream_t st, const char *what, int level = -1)
{
	if (!g_debug_sync) return;
	const cudaError_t e = cudaStreamSynchronize(st);
	if (e != cudaSuccess) {
		fprintf(stderr, "msb64 debug: %s (level %d): %s\n", what, level, cudaGetErrorString(e));
		abort();
	}
}

// Digit widths, most significant first (the role of schedule_passes, msb_64.c:1334), for
// keys of which only the low `width` bits vary (64 when nothing is known about the keys).
// *head (optional): the number of leading levels uniform keys need; the levels after them (the
// "tail") are TAIL_BITS wide each and run inside one cooperative launch (msb64_tail.cuh).
std::vector<int> make_schedule(uint64_t n, int width = 64, int *head = nullptr)
{
	if (width == 64 && !g_schedule_override.empty()) {
		if (head) *head = int(g_schedule_override.size());     // an override runs level by level
		return g_schedule_override;
	}
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
	// half of the local sort's capacity: LOCAL_CAP = 4096 -> buckets of 2^11 pairs
	constexpr int leaf_log = (LOCAL_CAP >= 4096 ? 12 : LOCAL_CAP >= 2048 ? 11 : 10) - 1;
	int need = log_n > leaf_log ? log_n - leaf_log : 0;
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
	// the rest of the key (only skewed inputs get here): 7-bit digits down to bit 0.  Dead
	// digits cost next to nothing (the plan kernel moves a segment down to its highest
	// differing bit), so the digits that do get used should be narrow enough for full-speed
	// scatter passes.  The last digit may reach below bit 0 / above `width`: a digit's position
	// is clamped at bit 0 and its surplus high bits were consumed by the level above.
	if (s.empty()) {
		s.push_back(TAIL_BITS);
		used += TAIL_BITS;
	}
	if (head) *head = int(s.size());
	for (int rest = width - used; rest > 0; rest -= TAIL_BITS) s.push_back(TAIL_BITS);
	return s;
}

// Level-0 digit of a sort whose keys are known to lie in [lo, hi]: (key >> shift0) - origin0
// with the schedule made for the bits that actually vary.  Returns the schedule; the
// digits below level 0 are plain bit fields under shift0.
struct RangePlan {
	std::vector<int> sched;
	int head;               // levels [head, sched.size()) are the tail (make_schedule)
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
		r.sched = make_schedule(n, width, &r.head);
		const int bits0 = r.sched[0];
		r.shift0 = width > bits0 ? width - bits0 : 0;
		r.origin0 = lo >> r.shift0;
		// the digit of the largest key must fit: (hi >> shift0) - origin0 < 2^bits0
		if (width >= 64 || ((hi >> r.shift0) - r.origin0) < (1ull << bits0)) break;
	}
	return r;
}

// ------------------------------------------------------------------ device state
// One record per CUDA device: kernel attributes, occupancy and cached allocations belong to
// the device that was current when they were made, so a process that drives several GPUs
// (the multi-device sort(), msb64_shard.cuh) gets one of each per device.
constexpr int MAX_DEVICES = 64;
struct Device {
	bool ready = false;
	int index = -1;
	int sms = 0;
	int hist_blocks[MAX_BITS + 1] = {0};      // resident blocks per SM, by digit width
	int fused_blocks[MAX_BITS + 1][FUSE_MAX_BITS + 1] = {{0}};   // same for the fused (two-level) histogram, by both widths
	int scatter_blocks[MAX_BITS + 1] = {0};
	int local_blocks = 0, packed_blocks = 0, tail_blocks = 0;
	bool route_configured = false;
	// cached allocations (grow-only)
	void *ws = nullptr;
	size_t ws_bytes = 0;
	uint64_t *dkeys = nullptr, *drids = nullptr;
	size_t dcap = 0;
	unsigned long long *scratch = nullptr;    // 128 words: sort()'s node boundaries [0, 64), check() sums [64, 67)
	// Sticky status of asynchronous sorts: the last kernel of a sort copies a non-zero
	// Control.error into this page-locked word, the host looks at it when the device is used
	// next (and in msb64_b200_last_status), so an untimed msb64_b200_sort_device whose work lists
	// overflowed cannot pass unnoticed.
	uint32_t *status_h = nullptr, *status_d = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4 * MAX_LEVELS + 16];
	bool events = false;
	// last sort
	uint64_t level_us[MAX_LEVELS][3] = {{0}};   // histogram, plan, scatter per level (timed sorts)
	int last_levels = 0;
	Control *last_ctl = nullptr;
	cudaStream_t last_stream = nullptr;
} g_devs[MAX_DEVICES];

template <int BITS>
int setup_bits(Device &D)
{
	using H = HistCfg<BITS, 256>;
	using S = ScatterCfg<BITS, SCATTER_THREADS>;
	CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<BITS, 256, false>,
				      cudaFuncAttributeMaxDynamicSharedMemorySize, int(H::SMEM)));
	if (BITS < FUSE_MAX_BITS - 3) {
		CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<BITS, 256, true>,
					      cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(H::SMEM + (size_t(H::NB + 32) << (FUSE_MAX_BITS - BITS)) * 4)));
		// the table of level-1 counts grows with the level-1 width: so does the block's footprint
		for (int fb = 4; BITS + fb <= FUSE_MAX_BITS; ++fb)
			CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
				&D.fused_blocks[BITS][fb], histogram_kernel<BITS, 256, true>, 256,
				H::SMEM + (size_t(H::NB + 32) << fb) * 4));
	}
	CUDA_TRY(cudaFuncSetAttribute(scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB>,
				      cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
		&D.hist_blocks[BITS], histogram_kernel<BITS, 256, false>, 256, H::SMEM));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
		&D.scatter_blocks[BITS], scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB>, SCATTER_THREADS, S::SMEM));
	if (D.hist_blocks[BITS] < 1 || D.scatter_blocks[BITS] < 1)
		return fail(MSB64_ERR_CUDA, "kernel does not fit on an SM%s");
	return MSB64_OK;
}

// The record of the CUDA device that is current on this thread, initialised on first use.
int device_get(Device **out)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(MSB64_ERR_CUDA, "no CUDA device: %s",
			    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	if (dev < 0 || dev >= MAX_DEVICES) return fail(MSB64_ERR_CUDA, "device index out of range%s");
	Device &D = g_devs[dev];
	*out = &D;
	if (D.ready) return MSB64_OK;
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
	if (prop.major < 10)
		return fail(MSB64_ERR_CUDA, "device %s is not sm_100 class", prop.name);
	D.index = dev;
	D.sms = prop.multiProcessorCount;
	int rc;
	if ((rc = setup_bits<4>(D)) || (rc = setup_bits<5>(D)) || (rc = setup_bits<6>(D)) ||
	    (rc = setup_bits<7>(D)) || (rc = setup_bits<8>(D)) || (rc = setup_bits<9>(D)) ||
	    (rc = setup_bits<10>(D)) || (rc = setup_bits<11>(D)))
		return rc;
	CUDA_TRY(cudaFuncSetAttribute(local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				      int(LOCAL_SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.local_blocks, local_sort_kernel,
							       LOCAL_THREADS, LOCAL_SMEM));
	if (D.local_blocks < 1) return fail(MSB64_ERR_CUDA, "local sort does not fit on an SM%s");
	CUDA_TRY(cudaFuncSetAttribute(local_sort_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
				      int(PACKED_SMEM)));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.packed_blocks, local_sort_packed_kernel,
							       LOCAL_THREADS, PACKED_SMEM));
	if (D.packed_blocks < 1) return fail(MSB64_ERR_CUDA, "packed local sort does not fit on an SM%s");
	CUDA_TRY(cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tail_smem())));
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.tail_blocks, tail_kernel, TAIL_THREADS, tail_smem()));
	int coop = 0;
	CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	if (D.tail_blocks < 1 || !coop) return fail(MSB64_ERR_CUDA, "tail kernel: no cooperative launch / does not fit on an SM%s");
	CUDA_TRY(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaMalloc(&D.scratch, 128 * sizeof(unsigned long long)));
	CUDA_TRY(cudaHostAlloc(&D.status_h, sizeof(uint32_t), cudaHostAllocMapped));
	*D.status_h = 0;
	CUDA_TRY(cudaHostGetDevicePointer(&D.status_d, D.status_h, 0));
	D.ready = true;
	return MSB64_OK;
}

#define DEVICE_OR_RETURN()                          \
	Device *dev_ = nullptr;                    \
	{                                          \
		const int rc_ = device_get(&dev_); \
		if (rc_) return rc_;               \
	}                                          \
	Device &D = *dev_

// Non-zero: an earlier asynchronous sort on this device reported an internal error (cleared).
uint32_t take_status(Device &D)
{
	const uint32_t e = *static_cast<volatile uint32_t *>(D.status_h);
	if (e) *static_cast<volatile uint32_t *>(D.status_h) = 0;
	return e;
}

int ensure_events(Device &D)
{
	if (D.events) return MSB64_OK;
	for (auto &ev : D.ev) CUDA_TRY(cudaEventCreate(&ev));
	D.events = true;
	return MSB64_OK;
}

// ------------------------------------------------------------------ workspace
struct Layout {
	size_t keys_b, rids_b, segs[2], tiles[2], hist[2], segbits[2], units, copies, ctl, fused, total;
	uint32_t max_segs, max_tiles, max_units, max_copies;
};

// Sized for the worst case over every schedule a sort of n pairs may get (full width, a key
// range of any width, an override): the histogram rows take the widest digit any default
// schedule uses (8 bits; MAX_BITS under an override) and the unit list the deepest recursion,
// so that msb64_b200_workspace_bytes(n) is enough for msb64_b200_sort_device_range too.
Layout make_layout(uint64_t n)
{
	Layout L;
	const int maxbits = g_schedule_override.empty() ? 8 : MAX_BITS;
	const uint64_t levels = MAX_LEVELS;
	L.max_segs = uint32_t(n / UNIT_CAP + 2);
	L.max_tiles = uint32_t(n / TILE + 2 * uint64_t(L.max_segs) + 2);
	L.max_units = uint32_t(2 * (n / UNIT_CAP) + 2 * levels * L.max_segs + 16);
	L.max_copies = uint32_t(n / COPY_TILE + L.max_segs + 2);
	size_t at = 0;
	auto take = [&](size_t bytes) { size_t o = at; at = align_up(at + bytes); return o; };
	L.keys_b = take((n + 2) * 8);                 // + an odd first element and the bulk window's even rounding
	L.rids_b = take((n + 2) * 8);
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
void launch_level(Device &D, const Ctx &c, int level, int shift0, uint32_t origin, int next_bits,
		  cudaStream_t st, cudaEvent_t *ev)
{
	using H = HistCfg<BITS, 256>;
	using S = ScatterCfg<BITS, SCATTER_THREADS>;
	if (ev) cudaEventRecord(ev[0], st);
	// level 0 is one segment: its histogram pass also counts the level-1 digits per bin
	// (32 KiB of shared counters), and level 1 needs no histogram pass
	const bool fuse = level == 0 && next_bits > 0 && BITS + next_bits <= FUSE_MAX_BITS &&
			  BITS < FUSE_MAX_BITS - 3 && D.fused_blocks[BITS][next_bits] > 0 && shift0 >= next_bits && !g_no_fuse;
	if (fuse) {
		const size_t smem = H::SMEM + (size_t(H::NB + 32) << next_bits) * 4;
		histogram_kernel<BITS, 256, true><<<D.sms * D.fused_blocks[BITS][next_bits], 256, smem, st>>>(
			c, level, origin, next_bits);
	} else {
		histogram_kernel<BITS, 256, false><<<D.sms * D.hist_blocks[BITS], 256, H::SMEM, st>>>(
			c, level, origin, 0);
	}
	debug_sync(st, "histogram_kernel", level);
	if (ev) cudaEventRecord(ev[1], st);
	plan_kernel<<<D.sms * 4, PLAN_THREADS, 0, st>>>(c, level, BITS, next_bits, fuse);
	debug_sync(st, "plan_kernel", level);
	if (ev) cudaEventRecord(ev[2], st);
	scatter_kernel<BITS, SCATTER_THREADS, SCATTER_MINB><<<D.sms * D.scatter_blocks[BITS], SCATTER_THREADS, S::SMEM, st>>>(c, level, origin);
	debug_sync(st, "scatter_kernel", level);
	if (ev) cudaEventRecord(ev[3], st);
	g_launches += 3;
}

int sort_device_locked(uint64_t *d_keys, uint64_t *d_rids, uint64_t n, void *workspace,
		       size_t workspace_bytes, cudaStream_t st, uint64_t *phase_us,
		       uint64_t key_lo = 0, uint64_t key_hi = ~0ull, uint64_t begin = 0)
{
	Device *dev = nullptr;
	int rc = device_get(&dev);
	if (rc) return rc;
	Device &D = *dev;
	if (take_status(D))
		return fail(MSB64_ERR_INTERNAL, "an earlier asynchronous sort on this device overflowed its work lists%s");
	if (n > MSB64_MAX_PAIRS || begin + n > MSB64_MAX_PAIRS)
		return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	if (phase_us) memset(phase_us, 0, MSB64_PHASE_COUNT * sizeof(uint64_t));
	if (n < 2) return MSB64_OK;
	if (!d_keys || !d_rids || (uintptr_t(d_keys) & 15) || (uintptr_t(d_rids) & 15))
		return fail(MSB64_ERR_ARG, "device arrays must be non-NULL and 16-byte aligned%s");

	const RangePlan rp = plan_range(n, key_lo, key_hi);
	const std::vector<int> &sched = rp.sched;
	const Layout L = make_layout(n);
	if (!workspace) {
		if (D.ws_bytes < L.total) {
			if (D.ws) cudaFree(D.ws);
			D.ws = nullptr;
			D.ws_bytes = 0;
			CUDA_TRY(cudaMalloc(&D.ws, L.total));
			D.ws_bytes = L.total;
		}
		workspace = D.ws;
	} else if (workspace_bytes < L.total || (uintptr_t(workspace) & 255)) {
		return fail(MSB64_ERR_NOMEM, "workspace too small or not 256-byte aligned%s");
	}
	char *w = static_cast<char *>(workspace);
	Ctx c;
	c.keys[0] = d_keys;
	c.rids[0] = d_rids;
	// the scratch copy holds elements [begin & ~1, begin + n) only: its base is shifted so that
	// both buffers are indexed by the same element numbers
	c.keys[1] = reinterpret_cast<uint64_t *>(w + L.keys_b) - (begin & ~1ull);
	c.rids[1] = reinterpret_cast<uint64_t *>(w + L.rids_b) - (begin & ~1ull);
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
	c.status = D.status_d;
	c.begin = uint32_t(begin);
	c.n = uint32_t(n);
	c.end = uint32_t(begin + n);
	c.max_segs = L.max_segs;
	c.max_tiles = L.max_tiles;
	c.max_units = L.max_units;
	c.max_copies = L.max_copies;

	cudaEvent_t *ev = nullptr;
	if (phase_us) {
		if ((rc = ensure_events(D))) return rc;
		ev = D.ev;
	}
	const int levels = int(sched.size());
	// levels [0, head) one by one, levels [head, levels) -- all TAIL_BITS wide, reached by skewed
	// inputs only -- inside one cooperative launch that leaves at the first empty level
	int head = std::min(std::max(rp.head, 1), levels);
	if (g_tail_from > 0 && g_tail_from < levels) {
		head = g_tail_from;
		for (int l = head; l < levels; ++l)
			if (sched[l] != TAIL_BITS) return fail(MSB64_ERR_ARG, "MSB64_TAIL_FROM: tail digits must be 7 bits wide%s");
	}
	if (ev) cudaEventRecord(ev[0], st);
	init_kernel<<<D.sms, 256, 0, st>>>(c, sched[0], rp.shift0);
	debug_sync(st, "init_kernel");
	g_launches += 1;
	cudaEvent_t *tail = ev ? ev + 1 + 4 * head : nullptr;
	if (n > LOCAL_CAP) {
		// the position of a segment's digit travels with the segment (msb64_plan.cuh); the
		// host only says where level 0 starts
		for (int l = 0; l < head; ++l) {
			const int bits = sched[l];
			const int shift = rp.shift0;
			const uint32_t origin = l == 0 ? uint32_t(rp.origin0) : 0u;
			const int next_bits = l + 1 < levels ? sched[l + 1] : 0;
			cudaEvent_t *lev = ev ? ev + 1 + 4 * l : nullptr;
			switch (bits) {
			case 4: launch_level<4>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 5: launch_level<5>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 6: launch_level<6>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 7: launch_level<7>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 8: launch_level<8>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 9: launch_level<9>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 10: launch_level<10>(D, c, l, shift, origin, next_bits, st, lev); break;
			case 11: launch_level<11>(D, c, l, shift, origin, next_bits, st, lev); break;
			default: return fail(MSB64_ERR_ARG, "digit width outside 4..11%s");
			}
		}
		if (tail) cudaEventRecord(tail[0], st);
		if (head < levels) {
			Ctx cc = c;
			int first = head, last = levels - 1;
			void *args[] = {&cc, &first, &last};
			CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(tail_kernel), dim3(D.sms * D.tail_blocks),
							     dim3(TAIL_THREADS), args, tail_smem(), st));
			debug_sync(st, "tail_kernel");
			g_launches += 1;
		}
	} else if (tail) {
		cudaEventRecord(tail[0], st);
	}
	if (tail) cudaEventRecord(tail[1], st);
	// units whose keys leave room for a slot number in one word take the packed path, the rest
	// (small arrays, very deep levels never) the general one; an empty list costs a launch
	local_sort_packed_kernel<<<D.sms * D.packed_blocks, LOCAL_THREADS, PACKED_SMEM, st>>>(
		c, rp.origin0 << rp.shift0);
	debug_sync(st, "local_sort_packed_kernel");
	local_sort_kernel<<<D.sms * D.local_blocks, LOCAL_THREADS, LOCAL_SMEM, st>>>(
		c, rp.origin0 << rp.shift0);
	debug_sync(st, "local_sort_kernel");
	g_launches += 1;
	if (tail) cudaEventRecord(tail[2], st);
	copy_kernel<<<D.sms * 8, 256, 0, st>>>(c);
	debug_sync(st, "copy_kernel");
	if (tail) cudaEventRecord(tail[3], st);
	g_launches += 2;
	CUDA_TRY(cudaGetLastError());
	D.last_ctl = c.ctl;
	D.last_stream = st;

	if (phase_us) {
		CUDA_TRY(cudaStreamSynchronize(st));
		auto us = [&](cudaEvent_t a, cudaEvent_t b) {
			float ms = 0;
			cudaEventElapsedTime(&ms, a, b);
			return uint64_t(ms * 1000.0f + 0.5f);
		};
		phase_us[MSB64_PHASE_PLAN] += us(ev[0], ev[1]);
		if (n > LOCAL_CAP)
			for (int l = 0; l < head; ++l) {
				cudaEvent_t *lev = ev + 1 + 4 * l;
				D.level_us[l][0] = us(lev[0], lev[1]);
				D.level_us[l][1] = us(lev[1], lev[2]);
				D.level_us[l][2] = us(lev[2], lev[3]);
				phase_us[MSB64_PHASE_HISTOGRAM] += D.level_us[l][0];
				phase_us[MSB64_PHASE_PLAN] += D.level_us[l][1];
				phase_us[MSB64_PHASE_SCATTER] += D.level_us[l][2];
			}
		D.last_levels = n > LOCAL_CAP ? head : 0;
		phase_us[MSB64_PHASE_TAIL] = us(tail[0], tail[1]);
		phase_us[MSB64_PHASE_LOCAL] = us(tail[1], tail[2]);
		phase_us[MSB64_PHASE_COPY] = us(tail[2], tail[3]);
		if (take_status(D)) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
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
		case 5: {       // clustered: ~3000 keys share their 52 high bits and differ in the low 2
			const uint64_t clusters = n / 3000 + 1;
			k = (mix64((x % clusters) + 0x1234567ull) & ~0xfffull) | (mix64(x) & 3ull);
			break;
		}
		case 6: k = i == n / 2 ? (1ull << 63) : (x & 0xfffull); break;       // one outlier, the rest 12 bits
		case 7: {       // zipf-like (exponent 1.3): value = floor(u^(-1/0.3)), spread by a multiplier
			const double u = (double(x >> 11) + 1.0) * (1.0 / 9007199254740993.0);
			const double z = floor(pow(u, -1.0 / 0.3));
			k = (z >= 9.2e18 ? 0x7fffffffffffffffull : uint64_t(z)) * 0x9e3779b97f4a7c15ull;
			break;
		}
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

int ensure_device_arrays(Device &D, uint64_t n)
{
	if (D.dcap >= n) return MSB64_OK;
	if (D.dkeys) cudaFree(D.dkeys);
	if (D.drids) cudaFree(D.drids);
	D.dkeys = D.drids = nullptr;
	D.dcap = 0;
	CUDA_TRY(cudaMalloc(&D.dkeys, align_up(n * 8)));
	CUDA_TRY(cudaMalloc(&D.drids, align_up(n * 8)));
	D.dcap = n;
	return MSB64_OK;
}

int digit_histogram_locked(const uint64_t *d_keys, uint64_t n, int shift, int bits, uint64_t origin,
			   uint64_t *d_hist, uint64_t *d_minmax, cudaStream_t st)
{
	DEVICE_OR_RETURN();
	if (bits < 1 || bits > ROUTE_MAX_BITS || shift < 0 || shift >= 64 || !d_hist)
		return fail(MSB64_ERR_ARG, "digit_histogram: bad shift/bits%s");
	if (n > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(uint64_t) << bits, st));
	if (d_minmax) {
		CUDA_TRY(cudaMemsetAsync(d_minmax, 0xff, sizeof(uint64_t), st));
		CUDA_TRY(cudaMemsetAsync(d_minmax + 1, 0, sizeof(uint64_t), st));
	}
	if (n) {
		digit_histogram_kernel<<<D.sms * 8, 256, sizeof(uint32_t) << bits, st>>>(
			d_keys, n, shift, bits, uint32_t(origin), reinterpret_cast<unsigned long long *>(d_hist),
			reinterpret_cast<unsigned long long *>(d_minmax));
		g_launches += 1;
	}
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

const char *kPhaseNames[] = {
	"Copy to device time:      ", "Histogram time:           ", "Plan time:                ",
	"Scatter time:             ", "Local sort time:          ", "Copy home time:           ",
	"Tail levels time:         ", "Copy to host time:        ",
};

int sort_host_locked(uint64_t **keys, uint64_t **rids, uint64_t *size, int numa, double fudge,
		     char **description, uint64_t *times)
{
	Device *dev = nullptr;
	int rc = device_get(&dev);
	if (rc) return rc;
	Device &D = *dev;
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
	if ((rc = ensure_device_arrays(D, total))) return rc;
	if ((rc = ensure_events(D))) return rc;
	cudaStream_t st = D.stream;
	cudaEvent_t *ev = D.ev + (4 * MAX_LEVELS + 8);     // spare events past the per-level ones

	CUDA_TRY(cudaEventRecord(ev[0], st));
	uint64_t at = 0;
	for (int n = 0; n < numa; ++n) {
		if (!size[n]) continue;
		CUDA_TRY(cudaMemcpyAsync(D.dkeys + at, keys[n], size[n] * 8, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(D.drids + at, rids[n], size[n] * 8, cudaMemcpyHostToDevice, st));
		at += size[n];
	}
	CUDA_TRY(cudaEventRecord(ev[1], st));
	uint64_t phase[MSB64_PHASE_COUNT];
	rc = sort_device_locked(D.dkeys, D.drids, total, nullptr, 0, st, times ? phase : nullptr);
	if (rc) return rc;

	// node boundaries: exact quantiles, equal keys never split (msb_64.c:1596-1606)
	std::vector<uint64_t> bounds(numa, total);
	if (numa > 1) {
		uint64_t *d_bounds = reinterpret_cast<uint64_t *>(D.scratch);
		boundary_kernel<<<1, 64, 0, st>>>(D.dkeys, total, numa, d_bounds);
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
			CUDA_TRY(cudaMemcpyAsync(keys[n], D.dkeys + prev, cnt * 8, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaMemcpyAsync(rids[n], D.drids + prev, cnt * 8, cudaMemcpyDeviceToHost, st));
		}
		size[n] = cnt;
		prev = bounds[n];
	}
	CUDA_TRY(cudaEventRecord(ev[3], st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (take_status(D)) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
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

#include "msb64_shard.cuh"

namespace {

// ------------------------------------------------------------------ sort() over several GPUs
// numa >= 2 arrays and as many usable devices: node n's pairs go to GPU n, the GPUs run the
// sharded sort (msb64_shard.cuh) over peer memory, and node n gets the n-th key range back --
// the reference's own contract across NUMA nodes (msb_64.c:2261-2275, 1596-1606, 2180).
// Host <-> device copies of the nodes run concurrently, one PCIe link each.
// MSB64_B200_VIRTUAL_SHARDS=1 lets several shards share a device (node n on device n mod
// count): the same code path on a box with a single GPU.
std::vector<msb64_b200_shard *> g_host_shards;
const char *kShardPhaseNames[] = {
	"Copy to device time:      ", "Histogram + plan time:    ", "Route + exchange + sort:  ",
	"Copy to host time:        ",
};

int host_shard_devices(int numa)
{
	int count = 0;
	if (numa < 2 || cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return 0;
	const char *v = getenv("MSB64_B200_VIRTUAL_SHARDS");
	if (count >= numa) return getenv("MSB64_B200_SINGLE_DEVICE") ? 0 : numa;
	return v && atoi(v) > 0 ? count : 0;
}

void drop_host_shards()
{
	for (auto *S : g_host_shards) shard_free(S);
	g_host_shards.clear();
}

int sort_host_sharded_locked(uint64_t **keys, uint64_t **rids, uint64_t *size, int numa, double fudge,
			     char **description, uint64_t *times, int ndev)
{
	using clk = std::chrono::steady_clock;
	if (!keys || !rids || !size || numa < 2 || numa > SHARD_MAX_WORLD || !(fudge >= 1.0))
		return fail(MSB64_ERR_ARG, "bad keys/rids/size/numa/fudge%s");
	uint64_t total = 0;
	for (int n = 0; n < numa; ++n) {
		if (size[n] && (!keys[n] || !rids[n])) return fail(MSB64_ERR_ARG, "NULL array%s");
		if ((uintptr_t(keys[n]) & 15) || (uintptr_t(rids[n]) & 15))
			return fail(MSB64_ERR_ARG, "arrays must be 16-byte aligned (msb_64.c:2272)%s");
		if (uint64_t(double(size[n]) * fudge) + 2 > MSB64_MAX_PAIRS)
			return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs per GPU%s");
		total += size[n];
	}
	if (description) description[0] = nullptr;
	if (total == 0) return MSB64_OK;
	int prev_dev = 0;
	CUDA_TRY(cudaGetDevice(&prev_dev));
	struct Restore {
		int dev;
		~Restore() { cudaSetDevice(dev); }
	} restore{prev_dev};

	// the shards of this process (kept between calls while they are large enough)
	bool reuse = int(g_host_shards.size()) == numa;
	for (int n = 0; reuse && n < numa; ++n) {
		const msb64_b200_shard *S = g_host_shards[n];
		reuse = S->device == n % ndev && S->capacity >= size[n] &&
			S->recv_cap >= uint64_t(double(size[n]) * fudge);
	}
	if (!reuse) {
		drop_host_shards();
		for (int n = 0; n < numa; ++n) {
			CUDA_TRY(cudaSetDevice(n % ndev));
			msb64_b200_shard *S = shard_create_locked(n, numa, size[n], fudge);
			if (!S) {
				drop_host_shards();
				return MSB64_ERR_NOMEM;
			}
			g_host_shards.push_back(S);
			if (cudaMalloc(&S->in_keys, (S->capacity + 2) * 8) != cudaSuccess ||
			    cudaMalloc(&S->in_rids, (S->capacity + 2) * 8) != cudaSuccess ||
			    cudaHostAlloc(&S->h_hist, SHARD_SLOTS * 8, cudaHostAllocDefault) != cudaSuccess ||
			    cudaStreamCreateWithFlags(&S->main, cudaStreamNonBlocking) != cudaSuccess) {
				snprintf(g_err, sizeof(g_err), "sort(): device memory for node %d: %s", n,
					 cudaGetErrorString(cudaGetLastError()));
				drop_host_shards();
				return MSB64_ERR_NOMEM;
			}
		}
		const int rc = shard_connect_local_locked(g_host_shards.data(), numa);
		if (rc) {
			drop_host_shards();
			return rc;
		}
	}
	auto &sh = g_host_shards;
	auto sync_all = [&]() -> int {
		for (int n = 0; n < numa; ++n) {
			CUDA_TRY(cudaSetDevice(sh[n]->device));
			CUDA_TRY(cudaStreamSynchronize(sh[n]->main));
		}
		return MSB64_OK;
	};
	int rc;
	const auto t0 = clk::now();
	for (int n = 0; n < numa; ++n) {
		CUDA_TRY(cudaSetDevice(sh[n]->device));
		if (!size[n]) continue;
		CUDA_TRY(cudaMemcpyAsync(sh[n]->in_keys, keys[n], size[n] * 8, cudaMemcpyHostToDevice, sh[n]->main));
		CUDA_TRY(cudaMemcpyAsync(sh[n]->in_rids, rids[n], size[n] * 8, cudaMemcpyHostToDevice, sh[n]->main));
	}
	if ((rc = sync_all())) return rc;
	const auto t1 = clk::now();

	// steps 1-3: histograms, "all-gather" through page-locked host memory, the plan
	std::vector<uint64_t> hists(size_t(numa) * SHARD_SLOTS), caps(numa);
	for (int n = 0; n < numa; ++n) caps[n] = std::min(sh[n]->recv_cap, uint64_t(double(size[n]) * fudge));   // msb_64.c:1574
	for (int round = 0;; ++round) {
		for (int n = 0; n < numa; ++n) {
			if (round) sh[n]->plan = sh[0]->plan;            // the window every shard histograms on
			if ((rc = shard_histogram_locked(*sh[n], sh[n]->in_keys, size[n], round == 0, sh[n]->main))) return rc;
			DeviceGuard guard(sh[n]->device);
			CUDA_TRY(cudaMemcpyAsync(sh[n]->h_hist, sh[n]->d_hist, SHARD_SLOTS * 8, cudaMemcpyDeviceToHost, sh[n]->main));
		}
		if ((rc = sync_all())) return rc;
		for (int n = 0; n < numa; ++n) memcpy(&hists[size_t(n) * SHARD_SLOTS], sh[n]->h_hist, SHARD_SLOTS * 8);
		rc = shard_plan(sh[0]->plan, hists.data(), numa, caps.data(), round == 0);
		if (rc == 1 && round == 0) continue;
		if (rc) return rc;
		for (int n = 1; n < numa; ++n) sh[n]->plan = sh[0]->plan;
		break;
	}
	const auto t2 = clk::now();

	// steps 4-6, interleaved over the shards: nothing a shard waits for is enqueued after the wait
	for (int n = 0; n < numa; ++n) {
		sh[n]->timed = false;
		if ((rc = shard_prepare_locked(*sh[n], sh[n]->in_keys, sh[n]->in_rids, size[n]))) return rc;
	}
	for (int n = 0; n < numa; ++n)
		if ((rc = shard_route_exchange_locked(*sh[n], sh[n]->in_keys, sh[n]->in_rids, size[n], sh[n]->main))) return rc;
	for (int n = 0; n < numa; ++n)
		if ((rc = shard_wait_sort_locked(*sh[n], sh[n]->main))) return rc;
	if ((rc = sync_all())) return rc;
	const auto t3 = clk::now();

	for (int n = 0; n < numa; ++n) {
		CUDA_TRY(cudaSetDevice(sh[n]->device));
		const uint64_t cnt = sh[n]->recv_total;
		if (cnt) {
			CUDA_TRY(cudaMemcpyAsync(keys[n], sh[n]->recv_keys, cnt * 8, cudaMemcpyDeviceToHost, sh[n]->main));
			CUDA_TRY(cudaMemcpyAsync(rids[n], sh[n]->recv_rids, cnt * 8, cudaMemcpyDeviceToHost, sh[n]->main));
		}
		size[n] = cnt;
	}
	if ((rc = sync_all())) return rc;
	const auto t4 = clk::now();
	for (int n = 0; n < numa; ++n) {
		Device *dev = nullptr;
		CUDA_TRY(cudaSetDevice(sh[n]->device));
		if ((rc = device_get(&dev))) return rc;
		if (take_status(*dev)) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
	}
	if (times && description) {
		auto us = [](clk::time_point a, clk::time_point b) {
			return uint64_t(std::chrono::duration_cast<std::chrono::microseconds>(b - a).count());
		};
		times[0] = us(t0, t1);
		times[1] = us(t1, t2);
		times[2] = us(t2, t3);
		times[3] = us(t3, t4);
		for (int p = 0; p < 4; ++p) description[p] = const_cast<char *>(kShardPhaseNames[p]);
		description[4] = nullptr;
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
	const int ndev = host_shard_devices(numa);
	if (ndev > 0) return sort_host_sharded_locked(keys, rids, size, numa, fudge, description, times, ndev);
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
	return make_layout(n).total;
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
	Device *dev = nullptr;
	if (device_get(&dev)) return 0;
	Device &D = *dev;
	if (!D.last_ctl) return 0;
	if (cudaStreamSynchronize(D.last_stream) != cudaSuccess) return 0;
	Control h;
	if (cudaMemcpy(&h, D.last_ctl, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
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
	Device *dev = nullptr;
	if (device_get(&dev)) return 0;
	Device &D = *dev;
	int k = 0;
	for (int l = 0; l < D.last_levels; ++l)
		for (int j = 0; j < 3 && k < cap; ++j) out[k++] = D.level_us[l][j];
	return k;
}

int msb64_b200_digit_histogram(const uint64_t *d_keys, uint64_t n, int shift, int bits, uint64_t origin,
				uint64_t *d_hist, uint64_t *d_minmax, void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	return digit_histogram_locked(d_keys, n, shift, bits, origin, d_hist, d_minmax, static_cast<cudaStream_t>(stream));
}

static int route_common(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, int shift, int bits,
			uint64_t origin, const uint8_t *d_bin_to_dest, int ndest, uint32_t *d_cursors, const RouteDst &dst,
			void *stream)
{
	DEVICE_OR_RETURN();
	if (bits < 1 || bits > ROUTE_MAX_BITS || shift < 0 || shift >= 64 || ndest < 1 ||
	    ndest > ROUTE_MAX_DEST || !d_bin_to_dest || !d_cursors)
		return fail(MSB64_ERR_ARG, "route: bad shift/bits/ndest%s");
	if (n > MSB64_MAX_PAIRS) return fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs%s");
	if (!n) return MSB64_OK;
	if (!D.route_configured) {
		CUDA_TRY(cudaFuncSetAttribute(route_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(RouteCfg<16>::SMEM)));
		CUDA_TRY(cudaFuncSetAttribute(route_kernel<ROUTE_MAX_DEST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(RouteCfg<ROUTE_MAX_DEST>::SMEM)));
		D.route_configured = true;
	}
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (ndest <= 16)
		route_kernel<16><<<D.sms * 3, ROUTE_THREADS, RouteCfg<16>::SMEM, st>>>(
			d_keys, d_rids, uint32_t(n), shift, bits, uint32_t(origin), d_bin_to_dest, ndest, d_cursors, dst);
	else
		route_kernel<ROUTE_MAX_DEST><<<D.sms * 2, ROUTE_THREADS, RouteCfg<ROUTE_MAX_DEST>::SMEM, st>>>(
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
	DEVICE_OR_RETURN();
	if (!n) return MSB64_OK;
	fill_kernel<<<D.sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_keys, d_rids, n, kind,
										seed, param);
	g_launches += 1;
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

int msb64_b200_check(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n, uint64_t *out,
		     void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	DEVICE_OR_RETURN();
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	unsigned long long *d_out = D.scratch + 64;        // the device's cached scratch words
	CUDA_TRY(cudaMemsetAsync(d_out, 0, 3 * sizeof(unsigned long long), st));
	if (n) {
		check_kernel<<<D.sms * 8, 256, 0, st>>>(d_keys, d_rids, n, d_out);
		g_launches += 1;
	}
	CUDA_TRY(cudaMemcpyAsync(out, d_out, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return MSB64_OK;
}

int msb64_b200_last_status(void *stream)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	DEVICE_OR_RETURN();
	CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
	if (take_status(D)) return fail(MSB64_ERR_INTERNAL, "device work list overflow%s");
	return MSB64_OK;
}

} // extern "C"
