// msb64_shard.cuh -- the sort sharded over several B200s (included by msb64_b200.cu; host code).
//
// The role of the reference's cross-NUMA-node phase -- sample, range histogram, partition
// into per-node ranges, every thread sorts its range (msb_64.c:239-351, 497-699, 1546-1606,
// 1674-2198) -- re-thought for GPUs on NVLink / NVSwitch.  One Shard per GPU (one process per
// GPU under torch.distributed, or all Shards in one process behind sort(), msb64_b200.cu):
//
//   1. histogram of the keys' top 12 bits (+ smallest and largest key)          8 B/key, HBM
//   2. the caller all-gathers the histograms (32 KiB per rank)
//   3. shard_plan(): every rank cuts the bin axis into world x subs BUCKETS of near-equal
//      count: destination d owns buckets [d*subs, (d+1)*subs), ascending key ranges.  Same
//      data, same cut on every rank; the table also yields every (source, bucket) count and
//      with them every offset on every GPU -- nothing else is exchanged.  If the top bits do
//      not separate the keys (a narrow value range) steps 1-2 run once more on a 13-bit
//      window placed on [global min, global max]
//   4. bucket_route_kernel: ONE local MSD pass at HBM speed groups the rank's pairs by bucket:
//      foreign buckets contiguous in a staging buffer, own buckets at their final place
//   5. exchange, sub-range by sub-range: the copy engines move bucket (d, s) into GPU d's
//      receive buffer over NVLink (peer memory: CUDA IPC mappings or same-process peer
//      access), two streams (keys, rids), and a one-thread-per-peer kernel raises the flag
//      word (source, s) in every peer's memory behind them
//   6. meanwhile the rank's main stream waits for the flags of sub-range 0 from all sources,
//      sorts it with the single-GPU sort told its key range (first digit relative to the
//      sub-range's lower end: no pass is spent on the bits the partition fixed), then
//      sub-range 1, ...: NVLink moves sub-range s+1.. while HBM sorts sub-range s.
//
// Afterwards rank r holds the r-th key range in ascending order, contiguous in its receive
// buffer (the contract of sort() across NUMA nodes, msb_64.c:2180).  A destination that would
// receive more than its capacity is an error on every rank (msb_64.c:1574-1578).
#pragma once

namespace {

constexpr int SHARD_MAX_WORLD = ROUTE_MAX_DEST;          // 64
constexpr int SHARD_MAX_BUCKETS = 256;                   // bucket ids are bytes
constexpr int SHARD_BITS = 12;                           // first-round digit: the keys' top 12 bits
constexpr int SHARD_SLOTS = (2 << SHARD_BITS) + 2;       // a histogram row: 2^13 counts (window round) + min + max
constexpr int SHARD_SUBS = 32;                           // most sub-ranges per destination (pipeline depth)
constexpr int SHARD_ROUTE_BUCKETS = 128;                 // most buckets the route pass is asked for

// Sub-ranges per destination.  More of them shorten the wait for the first one and make the
// sub-range sorts cheaper per pair (2^25 pairs need 14 bits = two 7-bit passes with a fused
// level-1 histogram; 2^26 pairs need 8 + 7: a slower 8-bit pass and a separate histogram pass),
// but every sort has a fixed cost (~0.1 ms) and the route pass slows down beyond 128 buckets
// (16-pair runs): 32 sub-ranges up to 4 GPUs, 16 on 8.
// Sub-ranges of every destination that the route pass stores straight into the destination's
// receive buffer (peer stores over NVLink, which has nothing else to do during the pass) instead
// of staging them for the copy engines: a quarter of the exchange leaves the GPU inside the
// route kernel at no extra HBM traffic, the first sub-ranges are complete when the pass ends
// (the sorts start at once), and the copy engines -- which only get ~370-480 GB/s while the
// sorts saturate HBM -- have less to move behind the sorts' back.
// MSB64_SHARD_DIRECT_EIGHTHS (developer switch): the fraction in eighths, 0..8.
constexpr int SHARD_DIRECT_EIGHTHS = 2;
inline int shard_direct(int world, int subs)
{
	static const int eighths = [] {
		const char *v = getenv("MSB64_SHARD_DIRECT_EIGHTHS");
		return v ? std::min(std::max(atoi(v), 0), 8) : SHARD_DIRECT_EIGHTHS;
	}();
	if (world < 2) return 0;
	return eighths ? std::max(1, subs * eighths / 8) : 0;
}

inline int shard_subs(int world)
{
	int subs = SHARD_SUBS;
	while (subs > 1 && world * subs > SHARD_ROUTE_BUCKETS) subs >>= 1;
	return subs;
}

// ------------------------------------------------------------------ the plan (host only)
struct ShardPlan {
	int world = 0, subs = 0, nbk = 0;
	int shift = 64 - SHARD_BITS, bits = SHARD_BITS;
	uint64_t origin = 0;
	std::vector<uint8_t> table;          // [2^bits] bin -> bucket
	std::vector<uint64_t> counts;        // [world][nbk] pairs source r holds of bucket b
	std::vector<uint64_t> total;         // [nbk] over all sources
	std::vector<int> first_bin, last_bin;   // [nbk] bins of a bucket (first > last: none)
	bool valid = false;

	uint64_t dest_total(int d) const
	{
		uint64_t t = 0;
		for (int s = 0; s < subs; ++s) t += total[d * subs + s];
		return t;
	}
	// element offset in destination d's receive buffer where source r's share of sub-range s starts
	uint64_t recv_offset(int d, int s, int r) const
	{
		uint64_t at = 0;
		for (int q = 0; q < s; ++q) at += total[d * subs + q];
		for (int q = 0; q < r; ++q) at += counts[size_t(q) * nbk + d * subs + s];
		return at;
	}
	void key_range(int b, uint64_t *lo, uint64_t *hi) const
	{
		const unsigned __int128 a = (unsigned __int128)(origin + uint64_t(first_bin[b])) << shift;
		const unsigned __int128 z = ((unsigned __int128)(origin + uint64_t(last_bin[b]) + 1) << shift) - 1;
		const unsigned __int128 top = ~0ull;
		*lo = uint64_t(a > top ? top : a);
		*hi = uint64_t(z > top ? top : z);
	}
};

// Cut `nb` bins with counts h[] into `parts` contiguous ranges of near-equal count: a bin goes
// to the part into which its midpoint falls (integer arithmetic: every rank gets the same cut).
void cut_bins(const uint64_t *h, int nb, int parts, uint64_t total, int *part_of)
{
	unsigned __int128 cum = 0;
	int prev = 0;
	for (int i = 0; i < nb; ++i) {
		int p = prev;
		if (total && h[i]) {
			const unsigned __int128 mid2 = 2 * cum + h[i];         // 2 x midpoint
			p = int(mid2 * unsigned(parts) / (2 * (unsigned __int128) total));
			if (p > parts - 1) p = parts - 1;
			if (p < prev) p = prev;
		}
		part_of[i] = p;
		prev = p;
		cum += h[i];
	}
}

// hists: [world][SHARD_SLOTS] rows of the digit (P.shift, P.bits, P.origin): counts, then min and
// max key at [2^bits] and [2^bits + 1].  recv_caps: [world] pairs every destination can take.
// Returns MSB64_OK (P filled), 1 (the digit was moved onto the keys' real span: histogram
// again and call again) or MSB64_ERR_CAPACITY.
int shard_plan(ShardPlan &P, const uint64_t *hists, int world, const uint64_t *recv_caps, bool may_retry)
{
	P.world = world;
	P.subs = shard_subs(world);
	P.nbk = world * P.subs;
	P.valid = false;
	const int nb = 1 << P.bits;
	std::vector<uint64_t> g(nb, 0);
	uint64_t total = 0, gmin = ~0ull, gmax = 0;
	bool any = false;
	for (int r = 0; r < world; ++r) {
		const uint64_t *row = hists + size_t(r) * SHARD_SLOTS;
		for (int i = 0; i < nb; ++i) g[i] += row[i];
		if (row[nb] <= row[nb + 1]) {                    // the rank holds keys
			gmin = row[nb] < gmin ? row[nb] : gmin;
			gmax = row[nb + 1] > gmax ? row[nb + 1] : gmax;
			any = true;
		}
	}
	for (int i = 0; i < nb; ++i) total += g[i];
	std::vector<int> dest(nb), sub(nb);
	cut_bins(g.data(), nb, world, total, dest.data());
	P.table.assign(nb, 0);
	for (int d = 0, i = 0; d < world; ++d) {
		int j = i;
		uint64_t dt = 0;
		while (j < nb && dest[j] == d) dt += g[j++];
		if (j > i) cut_bins(g.data() + i, j - i, P.subs, dt, sub.data() + i);
		for (int q = i; q < j; ++q) P.table[q] = uint8_t(d * P.subs + sub[q]);
		i = j;
	}
	P.counts.assign(size_t(world) * P.nbk, 0);
	P.total.assign(P.nbk, 0);
	P.first_bin.assign(P.nbk, nb);
	P.last_bin.assign(P.nbk, -1);
	for (int i = 0; i < nb; ++i) {
		const int b = P.table[i];
		if (P.first_bin[b] > i) P.first_bin[b] = i;
		P.last_bin[b] = i;
	}
	for (int r = 0; r < world; ++r) {
		const uint64_t *row = hists + size_t(r) * SHARD_SLOTS;
		for (int i = 0; i < nb; ++i) P.counts[size_t(r) * P.nbk + P.table[i]] += row[i];
	}
	for (int r = 0; r < world; ++r)
		for (int b = 0; b < P.nbk; ++b) P.total[b] += P.counts[size_t(r) * P.nbk + b];
	for (int d = 0; d < world; ++d)
		if (P.dest_total(d) > recv_caps[d]) {
			// the digit does not separate the keys: put a 13-bit window on their real span
			if (may_retry && any) {
				int width = 0;
				while (width < 64 && ((gmax - gmin) >> width)) ++width;
				const int shift2 = width > SHARD_BITS ? width - SHARD_BITS : 0;
				if (shift2 < P.shift) {
					P.shift = shift2;
					P.bits = SHARD_BITS + 1;             // digits 0 .. 2^12 inclusive
					P.origin = gmin >> shift2;
					return 1;
				}
			}
			snprintf(g_err, sizeof(g_err), "rank %d would receive %llu pairs, more than its capacity %llu",
				 d, (unsigned long long) P.dest_total(d), (unsigned long long) recv_caps[d]);
			return MSB64_ERR_CAPACITY;
		}
	P.valid = true;
	return MSB64_OK;
}

} // namespace

// ------------------------------------------------------------------ one GPU's share
struct msb64_b200_shard {
	int rank = 0, world = 1, device = 0;
	uint64_t capacity = 0, recv_cap = 0;
	// device memory owned by this shard
	uint64_t *recv_keys = nullptr, *recv_rids = nullptr;    // [recv_cap + 2]
	uint32_t *flags = nullptr;                              // [2 lanes][subs][world] epochs
	uint64_t *stage_keys = nullptr, *stage_rids = nullptr;  // [capacity]
	uint64_t *d_hist = nullptr;                             // [SHARD_SLOTS]
	uint8_t *d_table = nullptr;                             // [2^13]
	uint32_t *d_cursors = nullptr;                          // [SHARD_MAX_BUCKETS]
	void *ws = nullptr;
	size_t ws_bytes = 0;
	// the peers' memory as this process sees it ([rank] = own)
	uint64_t *peer_keys[SHARD_MAX_WORLD] = {nullptr}, *peer_rids[SHARD_MAX_WORLD] = {nullptr};
	uint32_t *peer_flags[SHARD_MAX_WORLD] = {nullptr};
	bool opened[SHARD_MAX_WORLD] = {false};                 // mapped through CUDA IPC (to be closed)
	bool connected = false;
	cudaStream_t xs[2] = {nullptr, nullptr};                // exchange streams: keys, rids
	cudaEvent_t ev_routed = nullptr, ev_x[2] = {nullptr, nullptr}, ev_done = nullptr;
	cudaEvent_t tev[6] = {nullptr};                         // timing: start, routed, first sub-range in, sorted, exchange end x2
	uint32_t epoch = 0;
	ShardPlan plan;
	uint64_t recv_total = 0, sent_total = 0, direct_total = 0;
	uint64_t key_lo = 0, key_hi = ~0ull;
	bool timed = false;
	// only when the shard is driven by the host-array sort() of this process (msb64_b200.cu)
	uint64_t *in_keys = nullptr, *in_rids = nullptr;        // [capacity + 2] the node's pairs on the device
	uint64_t *h_hist = nullptr;                             // [SHARD_SLOTS] page-locked
	cudaStream_t main = nullptr;
};

namespace {

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev)
	{
		cudaGetDevice(&prev);
		if (prev != dev) cudaSetDevice(dev);
		else prev = -1;
	}
	~DeviceGuard()
	{
		if (prev >= 0) cudaSetDevice(prev);
	}
};

void shard_free(msb64_b200_shard *S)
{
	DeviceGuard guard(S->device);
	cudaDeviceSynchronize();
	for (int r = 0; r < S->world; ++r)
		if (S->opened[r]) {
			cudaIpcCloseMemHandle(S->peer_keys[r]);
			cudaIpcCloseMemHandle(S->peer_rids[r]);
			cudaIpcCloseMemHandle(S->peer_flags[r]);
		}
	for (void *p : {(void *) S->recv_keys, (void *) S->recv_rids, (void *) S->flags, (void *) S->stage_keys,
			(void *) S->stage_rids, (void *) S->d_hist, (void *) S->d_table, (void *) S->d_cursors, S->ws,
			(void *) S->in_keys, (void *) S->in_rids})
		if (p) cudaFree(p);
	if (S->h_hist) cudaFreeHost(S->h_hist);
	if (S->main) cudaStreamDestroy(S->main);
	for (auto s : S->xs)
		if (s) cudaStreamDestroy(s);
	for (auto e : {S->ev_routed, S->ev_x[0], S->ev_x[1], S->ev_done})
		if (e) cudaEventDestroy(e);
	for (auto e : S->tev)
		if (e) cudaEventDestroy(e);
	delete S;
}

template <int NBK>
int launch_bucket_route(Device &D, const uint64_t *keys, const uint64_t *rids, uint64_t n, const ShardPlan &P,
			const uint8_t *d_table, uint32_t *d_cursors, const BucketOut &out, cudaStream_t st)
{
	static bool configured[MAX_DEVICES] = {false};
	if (!configured[D.index]) {
		CUDA_TRY(cudaFuncSetAttribute(bucket_route_kernel<NBK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					      int(BucketCfg<NBK>::SMEM)));
		configured[D.index] = true;
	}
	bucket_route_kernel<NBK><<<D.sms * BucketCfg<NBK>::MINB, BucketCfg<NBK>::THREADS, BucketCfg<NBK>::SMEM, st>>>(
		keys, rids, uint32_t(n), P.shift, (1u << P.bits) - 1, uint32_t(P.origin), d_table, d_cursors, out);
	g_launches += 1;
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

int shard_ensure_ws(msb64_b200_shard &S, size_t bytes)
{
	if (S.ws_bytes >= bytes) return MSB64_OK;
	if (S.ws) {
		CUDA_TRY(cudaDeviceSynchronize());
		cudaFree(S.ws);
	}
	S.ws = nullptr;
	S.ws_bytes = 0;
	CUDA_TRY(cudaMalloc(&S.ws, bytes));
	S.ws_bytes = bytes;
	return MSB64_OK;
}

// Steps 4-6 for one shard come in three calls so that a process driving several shards can
// interleave them (all allocations, then every shard's route + exchange, then every shard's
// wait + sort: a shard's waits only ever depend on work that has already been enqueued).
//
// prepare: checks, allocations (may synchronise the device: nothing of this step is in flight yet)
int shard_prepare_locked(msb64_b200_shard &S, const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n)
{
	DeviceGuard guard(S.device);
	const ShardPlan &P = S.plan;
	if (!P.valid || P.world != S.world) return fail(MSB64_ERR_ARG, "shard: no accepted plan%s");
	if (S.world > 1 && !S.connected) return fail(MSB64_ERR_ARG, "shard: peers not connected%s");
	if (n > S.capacity) return fail(MSB64_ERR_ARG, "shard: more pairs than the shard's capacity%s");
	const int subs = P.subs, nbk = P.nbk, me = S.rank;
	uint64_t mine = 0;
	for (int b = 0; b < nbk; ++b) mine += P.counts[size_t(me) * nbk + b];
	if (mine != n) return fail(MSB64_ERR_ARG, "shard: the plan was made for other data (count mismatch)%s");
	if (P.dest_total(me) > S.recv_cap) return fail(MSB64_ERR_CAPACITY, "shard: receive buffer too small%s");
	if (n && ((uintptr_t(d_keys) & 15) || (uintptr_t(d_rids) & 15) || !d_keys || !d_rids))
		return fail(MSB64_ERR_ARG, "device arrays must be non-NULL and 16-byte aligned%s");
	uint64_t largest = 0;
	for (int s = 0; s < subs; ++s) largest = std::max(largest, P.total[me * subs + s]);
	if (largest >= 2) {
		const size_t need = make_layout(largest).total;
		// some head room: the sub-ranges of the next call will not be exactly as long
		if (S.ws_bytes < need) return shard_ensure_ws(S, make_layout(largest + largest / 8 + 4096).total);
	}
	return MSB64_OK;
}

// route + exchange (steps 4, 5), enqueued without a host synchronisation; `st` is the shard's
// main stream
int shard_route_exchange_locked(msb64_b200_shard &S, const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n,
				cudaStream_t st)
{
	DeviceGuard guard(S.device);
	DEVICE_OR_RETURN();
	const ShardPlan &P = S.plan;
	const int W = S.world, subs = P.subs, nbk = P.nbk, me = S.rank;
	S.epoch += 1;
	S.recv_total = P.dest_total(me);
	{
		uint64_t lo = 0, hi = ~0ull, a, z;
		bool first = true;
		for (int s = 0; s < subs; ++s) {
			const int b = me * subs + s;
			if (P.first_bin[b] > P.last_bin[b]) continue;
			P.key_range(b, &a, &z);
			if (first) lo = a;
			hi = z;
			first = false;
		}
		S.key_lo = lo;
		S.key_hi = hi;
	}
	if (S.timed) CUDA_TRY(cudaEventRecord(S.tev[0], st));

	// 4. where every bucket of this source goes: the first `ndirect` sub-ranges of every peer
	//    straight into that peer's receive buffer (and all of the rank's own into its own), the
	//    other foreign buckets side by side in the staging buffer in the order they will travel
	//    (sub-range major)
	const int ndirect = shard_direct(W, subs);
	std::vector<uint32_t> cursors(SHARD_MAX_BUCKETS, 0);
	std::vector<uint64_t> stage_off(nbk, 0);
	uint64_t at = 0, direct = 0;
	for (int s = 0; s < subs; ++s)
		for (int j = 1; j < W; ++j) {
			const int d = (me + j) % W, b = d * subs + s;
			const uint64_t cnt = P.counts[size_t(me) * nbk + b];
			if (s < ndirect) {
				cursors[b] = uint32_t(P.recv_offset(d, s, me));
				direct += cnt;
			} else {
				stage_off[b] = at;
				cursors[b] = uint32_t(at);
				at += cnt;
			}
		}
	S.sent_total = at + direct;
	S.direct_total = direct;
	for (int s = 0; s < subs; ++s) cursors[me * subs + s] = uint32_t(P.recv_offset(me, s, me));
	CUDA_TRY(cudaMemcpyAsync(S.d_table, P.table.data(), P.table.size(), cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(S.d_cursors, cursors.data(), SHARD_MAX_BUCKETS * sizeof(uint32_t),
				 cudaMemcpyHostToDevice, st));
	FlagDst fd;
	for (int d = 0; d < ROUTE_MAX_DEST; ++d) fd.flag[d] = S.peer_flags[d < W ? d : me];
	if (n) {
		BucketOut out;
		for (int d = 0; d <= ROUTE_MAX_DEST; ++d) {
			out.keys[d] = d < W ? S.peer_keys[d] : S.stage_keys;
			out.rids[d] = d < W ? S.peer_rids[d] : S.stage_rids;
		}
		out.subs = uint32_t(subs);
		out.ndirect = uint32_t(ndirect);
		out.self = uint32_t(me);
		int rc;
		if (nbk <= 32) rc = launch_bucket_route<32>(D, d_keys, d_rids, n, P, S.d_table, S.d_cursors, out, st);
		else if (nbk <= 64) rc = launch_bucket_route<64>(D, d_keys, d_rids, n, P, S.d_table, S.d_cursors, out, st);
		else if (nbk <= 128) rc = launch_bucket_route<128>(D, d_keys, d_rids, n, P, S.d_table, S.d_cursors, out, st);
		else rc = launch_bucket_route<256>(D, d_keys, d_rids, n, P, S.d_table, S.d_cursors, out, st);
		if (rc) return rc;
	}
	if (W > 1) {
		// the directly stored sub-ranges of this source are complete at every peer
		shard_signal_direct_kernel<<<1, 256, 0, st>>>(fd, W, me, subs, ndirect, S.epoch);
		g_launches += 1;
	}
	CUDA_TRY(cudaEventRecord(S.ev_routed, st));
	if (S.timed) CUDA_TRY(cudaEventRecord(S.tev[1], st));

	// 5. exchange on two side streams: lane 0 carries keys, lane 1 rids
	if (W > 1) {
		for (int x = 0; x < 2; ++x) {
			cudaStream_t xs = S.xs[x];
			CUDA_TRY(cudaStreamWaitEvent(xs, S.ev_routed, 0));
			const uint64_t *stage = x ? S.stage_rids : S.stage_keys;
			for (int s = ndirect; s < subs; ++s) {
				for (int j = 1; j < W; ++j) {
					const int d = (me + j) % W, b = d * subs + s;
					const uint64_t cnt = P.counts[size_t(me) * nbk + b];
					if (!cnt) continue;
					uint64_t *dst = (x ? S.peer_rids[d] : S.peer_keys[d]) + P.recv_offset(d, s, me);
					CUDA_TRY(cudaMemcpyAsync(dst, stage + stage_off[b], cnt * 8, cudaMemcpyDefault, xs));
				}
				shard_signal_kernel<<<1, 64, 0, xs>>>(fd, W, me, uint32_t((x * subs + s) * W + me), S.epoch);
				g_launches += 1;
			}
			CUDA_TRY(cudaEventRecord(S.ev_x[x], xs));
			if (S.timed) CUDA_TRY(cudaEventRecord(S.tev[4 + x], xs));
		}
	}
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

// 6. sub-range by sub-range: wait for every source's share, sort it in place
int shard_wait_sort_locked(msb64_b200_shard &S, cudaStream_t st)
{
	DeviceGuard guard(S.device);
	const ShardPlan &P = S.plan;
	const int W = S.world, subs = P.subs, me = S.rank;
	uint64_t off = 0;
	for (int s = 0; s < subs; ++s) {
		const int b = me * subs + s;
		const uint64_t cnt = P.total[b];
		if (W > 1) {
			shard_wait_kernel<<<1, 128, 0, st>>>(S.flags, W, me, subs, s, 2, S.epoch);
			g_launches += 1;
		}
		if (s == 0 && S.timed) CUDA_TRY(cudaEventRecord(S.tev[2], st));
		if (cnt >= 2) {
			uint64_t lo, hi;
			P.key_range(b, &lo, &hi);
			const int rc = sort_device_locked(S.recv_keys, S.recv_rids, cnt, S.ws, S.ws_bytes, st, nullptr, lo, hi, off);
			if (rc) return rc;
		}
		off += cnt;
	}
	if (S.timed) CUDA_TRY(cudaEventRecord(S.tev[3], st));
	// the step is over for this stream when the outgoing copies are, too (the staging buffer is free)
	if (W > 1)
		for (int x = 0; x < 2; ++x) CUDA_TRY(cudaStreamWaitEvent(st, S.ev_x[x], 0));
	CUDA_TRY(cudaEventRecord(S.ev_done, st));
	CUDA_TRY(cudaGetLastError());
	return MSB64_OK;
}

// Step 1 with the shard's current digit (reset: back to the top 12 bits first).
int shard_histogram_locked(msb64_b200_shard &S, const uint64_t *d_keys, uint64_t n, bool reset, cudaStream_t st)
{
	if (n > S.capacity) return fail(MSB64_ERR_ARG, "shard: more pairs than the shard's capacity%s");
	if (reset) {
		S.plan.shift = 64 - SHARD_BITS;
		S.plan.bits = SHARD_BITS;
		S.plan.origin = 0;
		S.plan.valid = false;
	}
	DeviceGuard guard(S.device);
	const int nb = 1 << S.plan.bits;
	return digit_histogram_locked(d_keys, n, S.plan.shift, S.plan.bits, S.plan.origin, S.d_hist, S.d_hist + nb, st);
}

msb64_b200_shard *shard_create_locked(int rank, int world, uint64_t capacity, double fudge)
{
	Device *dev = nullptr;
	if (device_get(&dev)) return nullptr;
	if (world < 1 || world > SHARD_MAX_WORLD || rank < 0 || rank >= world || !(fudge >= 1.0)) {
		fail(MSB64_ERR_ARG, "shard_create: bad rank/world/fudge%s");
		return nullptr;
	}
	const uint64_t recv_cap = uint64_t(double(capacity) * fudge) + 2;
	if (recv_cap > MSB64_MAX_PAIRS) {
		fail(MSB64_ERR_TOO_BIG, "more than MSB64_MAX_PAIRS pairs per GPU%s");
		return nullptr;
	}
	msb64_b200_shard *S = new msb64_b200_shard;
	S->rank = rank;
	S->world = world;
	S->device = dev->index;
	S->capacity = capacity;
	S->recv_cap = recv_cap;
	const int subs = shard_subs(world);
	const size_t nflags = size_t(2) * subs * world;
	bool ok = cudaMalloc(&S->recv_keys, (recv_cap + 2) * 8) == cudaSuccess &&
		  cudaMalloc(&S->recv_rids, (recv_cap + 2) * 8) == cudaSuccess &&
		  cudaMalloc(&S->flags, nflags * 4) == cudaSuccess &&
		  cudaMalloc(&S->d_hist, SHARD_SLOTS * 8) == cudaSuccess &&
		  cudaMalloc(&S->d_table, size_t(2) << SHARD_BITS) == cudaSuccess &&
		  cudaMalloc(&S->d_cursors, SHARD_MAX_BUCKETS * 4) == cudaSuccess;
	if (ok && world > 1)
		ok = cudaMalloc(&S->stage_keys, (capacity + 2) * 8) == cudaSuccess &&
		     cudaMalloc(&S->stage_rids, (capacity + 2) * 8) == cudaSuccess;
	ok = ok && cudaMemset(S->flags, 0, nflags * 4) == cudaSuccess;
	for (int x = 0; x < 2 && ok; ++x) {
		ok = cudaStreamCreateWithFlags(&S->xs[x], cudaStreamNonBlocking) == cudaSuccess &&
		     cudaEventCreateWithFlags(&S->ev_x[x], cudaEventDisableTiming) == cudaSuccess;
	}
	ok = ok && cudaEventCreateWithFlags(&S->ev_routed, cudaEventDisableTiming) == cudaSuccess &&
	     cudaEventCreateWithFlags(&S->ev_done, cudaEventDisableTiming) == cudaSuccess;
	for (auto &e : S->tev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
	if (!ok) {
		snprintf(g_err, sizeof(g_err), "shard_create: %s", cudaGetErrorString(cudaGetLastError()));
		shard_free(S);
		return nullptr;
	}
	S->peer_keys[rank] = S->recv_keys;
	S->peer_rids[rank] = S->recv_rids;
	S->peer_flags[rank] = S->flags;
	S->connected = world == 1;
	return S;
}


int shard_connect_local_locked(msb64_b200_shard *const *shards, int world)
{
	if (!shards || world < 1 || world > SHARD_MAX_WORLD) return fail(MSB64_ERR_ARG, "shard_connect_local: bad arguments%s");
	for (int a = 0; a < world; ++a) {
		msb64_b200_shard *S = shards[a];
		if (!S || S->world != world || S->rank != a) return fail(MSB64_ERR_ARG, "shard_connect_local: shard list does not match%s");
		DeviceGuard guard(S->device);
		for (int b = 0; b < world; ++b) {
			if (shards[b]->device != S->device) {
				int can = 0;
				CUDA_TRY(cudaDeviceCanAccessPeer(&can, S->device, shards[b]->device));
				if (!can) return fail(MSB64_ERR_CUDA, "no peer access between the devices%s");
				cudaError_t e = cudaDeviceEnablePeerAccess(shards[b]->device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
				cudaGetLastError();
			}
			S->peer_keys[b] = shards[b]->recv_keys;
			S->peer_rids[b] = shards[b]->recv_rids;
			S->peer_flags[b] = shards[b]->flags;
		}
		S->connected = true;
	}
	return MSB64_OK;
}


} // namespace

// =================================================================== C ABI, section 5
extern "C" {

int msb64_b200_shard_slots(void) { return SHARD_SLOTS; }
int msb64_b200_shard_subs(int world) { return world >= 1 && world <= SHARD_MAX_WORLD ? shard_subs(world) : 0; }

int msb64_b200_shard_plan_host(const uint64_t *hists, int world, const uint64_t *recv_caps, int may_retry,
			       int *shift, int *bits, uint64_t *origin, uint8_t *table, uint64_t *counts)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	if (!hists || !recv_caps || !shift || !bits || !origin || world < 1 || world > SHARD_MAX_WORLD ||
	    *bits < 1 || *bits > SHARD_BITS + 1 || *shift < 0 || *shift > 63)
		return fail(MSB64_ERR_ARG, "shard_plan_host: bad arguments%s");
	ShardPlan P;
	P.shift = *shift;
	P.bits = *bits;
	P.origin = *origin;
	const int rc = shard_plan(P, hists, world, recv_caps, may_retry != 0);
	*shift = P.shift;
	*bits = P.bits;
	*origin = P.origin;
	if (rc == MSB64_OK) {
		if (table) memcpy(table, P.table.data(), P.table.size());
		if (counts) memcpy(counts, P.counts.data(), P.counts.size() * sizeof(uint64_t));
	}
	return rc;
}

msb64_b200_shard *msb64_b200_shard_create(int rank, int world, uint64_t capacity, double fudge)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	return shard_create_locked(rank, world, capacity, fudge);
}

void msb64_b200_shard_destroy(msb64_b200_shard *S)
{
	if (!S) return;
	std::lock_guard<std::mutex> lock(g_mutex);
	shard_free(S);
}

int msb64_b200_shard_export(msb64_b200_shard *S, void *handles)
{
	static_assert(MSB64_SHARD_HANDLE_BYTES == 3 * sizeof(cudaIpcMemHandle_t), "handle size");
	if (!S || !handles) return fail(MSB64_ERR_ARG, "shard_export: NULL%s");
	DeviceGuard guard(S->device);
	cudaIpcMemHandle_t h[3];
	CUDA_TRY(cudaIpcGetMemHandle(&h[0], S->recv_keys));
	CUDA_TRY(cudaIpcGetMemHandle(&h[1], S->recv_rids));
	CUDA_TRY(cudaIpcGetMemHandle(&h[2], S->flags));
	memcpy(handles, h, sizeof(h));
	return MSB64_OK;
}

int msb64_b200_shard_connect_ipc(msb64_b200_shard *S, const void *all_handles)
{
	if (!S || !all_handles) return fail(MSB64_ERR_ARG, "shard_connect_ipc: NULL%s");
	std::lock_guard<std::mutex> lock(g_mutex);
	DeviceGuard guard(S->device);
	for (int r = 0; r < S->world; ++r) {
		if (r == S->rank) continue;
		cudaIpcMemHandle_t h[3];
		memcpy(h, static_cast<const char *>(all_handles) + size_t(r) * MSB64_SHARD_HANDLE_BYTES, sizeof(h));
		void *p[3] = {nullptr, nullptr, nullptr};
		for (int i = 0; i < 3; ++i) {
			cudaError_t e = cudaIpcOpenMemHandle(&p[i], h[i], cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				snprintf(g_err, sizeof(g_err), "cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(e));
				cudaGetLastError();
				for (int q = 0; q < i; ++q) cudaIpcCloseMemHandle(p[q]);
				return MSB64_ERR_CUDA;
			}
		}
		S->peer_keys[r] = static_cast<uint64_t *>(p[0]);
		S->peer_rids[r] = static_cast<uint64_t *>(p[1]);
		S->peer_flags[r] = static_cast<uint32_t *>(p[2]);
		S->opened[r] = true;
	}
	S->connected = true;
	return MSB64_OK;
}

int msb64_b200_shard_connect_local(msb64_b200_shard *const *shards, int world)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	return shard_connect_local_locked(shards, world);
}

// Step 1 with the shard's current digit (reset = 1: back to the top 12 bits first).
int msb64_b200_shard_histogram(msb64_b200_shard *S, const uint64_t *d_keys, uint64_t n, int reset, void *stream)
{
	if (!S) return fail(MSB64_ERR_ARG, "shard_histogram: NULL%s");
	std::lock_guard<std::mutex> lock(g_mutex);
	return shard_histogram_locked(*S, d_keys, n, reset != 0, static_cast<cudaStream_t>(stream));
}

uint64_t *msb64_b200_shard_hist(msb64_b200_shard *S) { return S ? S->d_hist : nullptr; }

int msb64_b200_shard_plan(msb64_b200_shard *S, const uint64_t *all_hists, const uint64_t *recv_caps, int may_retry)
{
	if (!S || !all_hists || !recv_caps) return fail(MSB64_ERR_ARG, "shard_plan: NULL%s");
	std::lock_guard<std::mutex> lock(g_mutex);
	return shard_plan(S->plan, all_hists, S->world, recv_caps, may_retry != 0);
}

int msb64_b200_shard_exchange_sort(msb64_b200_shard *S, const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n,
				   void *stream, int timed)
{
	if (!S) return fail(MSB64_ERR_ARG, "shard_exchange_sort: NULL%s");
	std::lock_guard<std::mutex> lock(g_mutex);
	S->timed = timed != 0;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	int rc = shard_prepare_locked(*S, d_keys, d_rids, n);
	if (!rc) rc = shard_route_exchange_locked(*S, d_keys, d_rids, n, st);
	if (!rc) rc = shard_wait_sort_locked(*S, st);
	return rc;
}

uint64_t msb64_b200_shard_count(const msb64_b200_shard *S) { return S ? S->recv_total : 0; }
uint64_t msb64_b200_shard_sent(const msb64_b200_shard *S) { return S ? S->sent_total : 0; }
uint64_t msb64_b200_shard_sent_direct(const msb64_b200_shard *S) { return S ? S->direct_total : 0; }
uint64_t msb64_b200_shard_recv_capacity(const msb64_b200_shard *S) { return S ? S->recv_cap : 0; }
uint64_t *msb64_b200_shard_keys(msb64_b200_shard *S) { return S ? S->recv_keys : nullptr; }
uint64_t *msb64_b200_shard_rids(msb64_b200_shard *S) { return S ? S->recv_rids : nullptr; }

int msb64_b200_shard_key_range(const msb64_b200_shard *S, uint64_t *key_lo, uint64_t *key_hi)
{
	if (!S || !key_lo || !key_hi) return fail(MSB64_ERR_ARG, "shard_key_range: NULL%s");
	*key_lo = S->key_lo;
	*key_hi = S->key_hi;
	return MSB64_OK;
}

// Device times (milliseconds) of the last msb64_b200_shard_exchange_sort(timed = 1); the caller
// has synchronised the stream.  ms[0] route pass, ms[1] wait for the first sub-range after the
// route, ms[2] sorting (first sub-range in -> last one sorted), ms[3] exchange (route done ->
// last outgoing copy done, the slower lane), ms[4] whole step.
int msb64_b200_shard_times(msb64_b200_shard *S, double *ms)
{
	if (!S || !ms) return fail(MSB64_ERR_ARG, "shard_times: NULL%s");
	DeviceGuard guard(S->device);
	float t = 0;
	CUDA_TRY(cudaEventElapsedTime(&t, S->tev[0], S->tev[1]));
	ms[0] = t;
	CUDA_TRY(cudaEventElapsedTime(&t, S->tev[1], S->tev[2]));
	ms[1] = t;
	CUDA_TRY(cudaEventElapsedTime(&t, S->tev[2], S->tev[3]));
	ms[2] = t;
	ms[3] = 0;
	if (S->world > 1)
		for (int x = 0; x < 2; ++x) {
			CUDA_TRY(cudaEventElapsedTime(&t, S->tev[1], S->tev[4 + x]));
			ms[3] = t > ms[3] ? t : ms[3];
		}
	CUDA_TRY(cudaEventElapsedTime(&t, S->tev[0], S->tev[3]));
	ms[4] = t;
	return MSB64_OK;
}

} // extern "C"
