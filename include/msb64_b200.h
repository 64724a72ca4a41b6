/* msb64_b200.h -- C ABI of the B200-native MSD radix sort for 64-bit key + 64-bit rid pairs.
 *
 * This is the drop-in boundary for the one hot path this library replaces: the
 * in-place MSD radix sort of the reference's msb_64.c.  Plain C, plain pointers
 * and sizes; no torch or CUDA types in any signature (a stream is passed as
 * void*).  The library is libmsb64_b200.so, built by
 * inplacemsdradixsort_b200/build.py with nvcc for sm_100a only.
 *
 * Section 1 re-exports the reference's own public interface, unchanged
 * (reference include/msb_64.h:36-40), so a program written against msb_64.h links
 * against this library instead of msb_64.c and gets the same arrays back sorted.
 * Section 2 is the same sort with explicit error codes and with device-resident
 * data (what a GPU pipeline binds).  Section 3 holds the helpers the reference
 * keeps beside sort() (generator, checker) in their device form.
 *
 * There is no CPU fallback anywhere behind these entry points: if no CUDA device
 * is usable every int-returning function returns MSB64_ERR_CUDA, and sort()
 * prints the error and aborts (the reference's own failure mode is assert()).
 */
#ifndef MSB64_B200_H_
#define MSB64_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- 1. drop-in
 *
 * sort(): replaces msb_64.c:2261-2430 (and everything it spawns: sort_thread,
 * msb_64.c:1477-2259).  Same argument meaning:
 *   keys[n], rids[n]  numa host arrays, 16-byte aligned (msb_64.c:2272-2275), each with
 *                     room for size[n] * fudge pairs (msb_64.c:1574-1578);
 *   size[n]           pairs held by node n on entry; on return the pairs node n
 *                     holds after the global sort (msb_64.c:2180): node n gets the
 *                     n-th contiguous key range, ascending, and
 *                     keys[n][size[n]-1] <= keys[n+1][0];
 *   threads           accepted for compatibility (the reference asserts 64,
 *                     msb_64.c:2266); not used, the GPU picks its own grid;
 *   numa              number of arrays (1..64);
 *   fudge             capacity factor of every array, >= 1.0;
 *   description/times NULL, or arrays of >= 16 entries that receive a
 *                     NULL-terminated list of phase names and their times in
 *                     microseconds (the reference fills 10 entries,
 *                     msb_64.c:2391-2401; the phases named here are this
 *                     implementation's own).
 * Where the reference picks the node boundaries from a random sample with an
 * uninitialised seed (thread_data_t.seed is never written), this library uses
 * the exact numa-quantiles of the sorted keys and, like the reference, never
 * separates equal keys across a node boundary (msb_64.c:1596-1606).
 * The order of rids among equal keys is unspecified (MSD radix sort is not
 * stable; the same holds for the reference).
 * Aborts (like the reference's asserts) on: misaligned arrays, a node that would
 * overflow size[n]*fudge, more than MSB64_MAX_PAIRS pairs, or a CUDA error.  */
void sort(uint64_t **keys, uint64_t **rids, uint64_t *size,
	  int threads, int numa, double fudge,
	  char **description, uint64_t *times);

/* mamalloc(): replaces msb_64.c:111-115.  64-byte aligned host memory, released
 * with free() exactly like the reference's. */
void *mamalloc(size_t size);

/* ------------------------------------------------------ 2. explicit-error API */

#define MSB64_OK            0
#define MSB64_ERR_CUDA     -1	/* no device / CUDA runtime error (see msb64_b200_last_error) */
#define MSB64_ERR_ARG      -2	/* NULL or misaligned pointer, bad numa, fudge < 1 */
#define MSB64_ERR_TOO_BIG  -3	/* more than MSB64_MAX_PAIRS pairs in one call */
#define MSB64_ERR_CAPACITY -4	/* a node would exceed size[n] * fudge */
#define MSB64_ERR_NOMEM    -5	/* device or workspace memory exhausted */
#define MSB64_ERR_INTERNAL -6	/* device-side work list overflow (a bug) */

/* Largest number of pairs one call sorts on one GPU (32-bit element indices). */
#define MSB64_MAX_PAIRS 0xFFFF0000ull

#define MSB64_MAX_PHASES 16

/* Same as sort() but returns an error code instead of aborting. */
int msb64_b200_sort(uint64_t **keys, uint64_t **rids, uint64_t *size,
		    int threads, int numa, double fudge,
		    char **description, uint64_t *times);

/* One host array pair, in place (the numa == 1 case of sort()). */
int msb64_b200_sort_host(uint64_t *keys, uint64_t *rids, uint64_t n);

/* Device-resident sort.  d_keys / d_rids: device pointers to n pairs, sorted in
 * place (the result always lands in these arrays).  workspace: device memory of
 * at least msb64_b200_workspace_bytes(n) bytes, 256-byte aligned, or NULL to let
 * the library allocate and cache one.  stream: a cudaStream_t passed as void*
 * (NULL = the default stream); the call only enqueues work and does not
 * synchronise, unless phase_us != NULL, in which case it synchronises and fills
 * phase_us[0..MSB64_PHASE_COUNT) with device times in microseconds.
 * This is local_radixsort's role (msb_64.c:1007-1035) for the whole array. */
size_t msb64_b200_workspace_bytes(uint64_t n);
int msb64_b200_sort_device(uint64_t *d_keys, uint64_t *d_rids, uint64_t n,
			   void *workspace, size_t workspace_bytes,
			   void *stream, uint64_t *phase_us);

/* The same sort when every key is known to lie in [key_lo, key_hi] (a range partition's
 * share, msb_64.c:1546-1606: after the reference's range partition a thread's keys lie
 * between two delimiters).  The digit schedule is made for the bits that vary inside the
 * range and the first digit is taken relative to key_lo, so a narrow range costs fewer
 * passes.  A key outside the range makes the result undefined (it is not checked). */
int msb64_b200_sort_device_range(uint64_t *d_keys, uint64_t *d_rids, uint64_t n,
				 void *workspace, size_t workspace_bytes,
				 void *stream, uint64_t *phase_us,
				 uint64_t key_lo, uint64_t key_hi);

/* Phases reported by msb64_b200_sort_device (indices into phase_us). */
#define MSB64_PHASE_HISTOGRAM 0	/* per-digit histogram kernels        (msb_64.c:701-738)  */
#define MSB64_PHASE_PLAN      1	/* bucket scans / work lists          (msb_64.c:1020-1034) */
#define MSB64_PHASE_SCATTER   2	/* partition kernels                  (msb_64.c:740-978)  */
#define MSB64_PHASE_LOCAL     3	/* small-bucket finish in shared mem  (msb_64.c:126-149, 980-1005) */
#define MSB64_PHASE_COPY      4	/* buckets finished in the scratch buffer copied home */
#define MSB64_PHASE_TAIL      5	/* histogram + plan + scatter of the levels below the depth uniform
				   keys need, one cooperative launch (skewed inputs only; msb_64.c:1007-1035) */
#define MSB64_PHASE_COUNT     6

/* Digit schedule: widths of the MSD digits, most significant first, covering all 64
 * bits (the role of schedule_passes, msb_64.c:1334-1400): they sum to at least 64, the
 * last digit starts above bit 0 and is clamped there.  Returns the number of levels
 * written to bits[] (<= 16).  msb64_b200_set_schedule overrides the default choice for
 * later calls with widths that sum to exactly 64 (count = 0 restores the default). */
int msb64_b200_get_schedule(uint64_t n, int *bits);
int msb64_b200_set_schedule(const int *bits, int count);
/* The schedule msb64_b200_sort_device_range uses for keys in [key_lo, key_hi]: digit widths
 * in bits[] (returns their number), position of the first digit in *shift0 and its origin
 * (key_lo >> shift0) in *origin0: first digit = (key >> shift0) - origin0; the digits below
 * are bit fields whose widths follow bits[] and whose positions start right under shift0
 * (on the device a segment whose keys agree on more bits is moved further down, see
 * DESIGN.md; the last digit may be wider than the key bits left: its surplus high bits were
 * consumed by the level above).  Needs no device. */
int msb64_b200_get_range_schedule(uint64_t n, uint64_t key_lo, uint64_t key_hi, int *bits,
				  int *shift0, uint64_t *origin0);

/* Tuning / introspection. */
int msb64_b200_device_count(void);
const char *msb64_b200_last_error(void);
/* Number of kernels launched by this library since load (for bench.py). */
uint64_t msb64_b200_launch_count(void);
/* Statistics of the last msb64_b200_sort_device call, read back from the device
 * (synchronises the stream): out[0..8) = segments per level 0..7 ... see DESIGN.md.
 * Returns the number of values written. */
int msb64_b200_last_stats(uint64_t *out, int cap);

/* The device sorts are asynchronous: a work-list overflow inside one (MSB64_ERR_INTERNAL, a
 * bug) is recorded in a page-locked status word and reported by the next call that uses the
 * device -- or by this function, which synchronises `stream` first and clears the status. */
int msb64_b200_last_status(void *stream);

/* Per-level device times of the last msb64_b200_sort_device call that passed
 * phase_us: out[3*l + 0..2] = histogram, plan, scatter microseconds of level l.
 * Returns the number of values written (3 * levels). */
int msb64_b200_last_level_times(uint64_t *out, int cap);

/* --------------------------------------------------------------- 3. helpers */

/* Page-locked host memory for fast host<->device copies (free with
 * msb64_b200_host_free, NOT free()). */
void *msb64_b200_host_alloc(size_t bytes);
void msb64_b200_host_free(void *p);

/* Device memory (so that non-CUDA callers can use the device API over the C ABI). */
void *msb64_b200_device_alloc(size_t bytes);
void msb64_b200_device_free(void *p);
int msb64_b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int msb64_b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int msb64_b200_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int msb64_b200_stream_sync(void *stream);

/* Synthetic inputs generated on the device (counter-based splitmix64; NOT the
 * reference's MT19937-64 of rand.c, which is sequential):
 *   kind 0: uniform 64-bit keys            kind 1: keys & mask (mask = param)
 *   kind 2: `param` distinct values        kind 3: ascending   kind 4: descending
 *   kind 5: clusters of ~3000 keys that differ in their low 2 bits only
 *   kind 6: 12-bit keys and one outlier    kind 7: zipf-like (exponent 1.3)
 * rids = element index when d_rids != NULL. */
int msb64_b200_fill(uint64_t *d_keys, uint64_t *d_rids, uint64_t n,
		    int kind, uint64_t seed, uint64_t param, void *stream);

/* Device form of check() (msb_64.c:2432-2505): returns in out[0] the number of
 * descents keys[i] > keys[i+1], out[1] the wrapping sum of keys (the reference's
 * checksum), out[2] an order-independent digest of the (key, rid) multiset.
 * Synchronises the stream. */
int msb64_b200_check(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n,
		     uint64_t *out, void *stream);

/* ------------------------------------------- 4. multi-GPU range partition steps
 *
 * The device halves of the cross-GPU range partition (the role of the reference's
 * sample / range_histogram / partition phase across NUMA nodes, msb_64.c:239-351,
 * 497-699, 1546-1606); the collectives between them are the caller's
 * (inplacemsdradixsort_b200/distributed.py uses NCCL through torch.distributed).
 *
 * Digits of these steps: digit(key) = ((key >> shift) - origin) & (2^bits - 1), bits <= 13.
 * With origin = 0 that is a plain bit field; a caller that knows the smallest key (below)
 * passes origin = smallest >> shift and a shift chosen for the keys' real span, so that
 * keys sharing a long prefix (only low bits significant) still spread over the bins.
 *
 * msb64_b200_digit_histogram: d_hist[0 .. 2^bits) = number of keys per digit (zeroed
 *   first).  d_minmax: NULL, or two words that receive the smallest and the largest key.
 * msb64_b200_route: groups the n pairs by destination = d_bin_to_dest[digit] (one byte
 *   per bin, ndest <= 64 destinations) into d_out_keys / d_out_rids; d_cursors[dest]
 *   must hold the first output slot of each destination (exclusive prefix of the
 *   per-destination counts) and is advanced by the kernel.  Order inside a
 *   destination is unspecified. */
int msb64_b200_digit_histogram(const uint64_t *d_keys, uint64_t n, int shift, int bits,
			       uint64_t origin, uint64_t *d_hist, uint64_t *d_minmax, void *stream);
int msb64_b200_route(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n,
		     int shift, int bits, uint64_t origin, const uint8_t *d_bin_to_dest, int ndest,
		     uint32_t *d_cursors, uint64_t *d_out_keys, uint64_t *d_out_rids,
		     void *stream);

/* Fused route + exchange over peer memory: the same kernel, but destination d has its own
 * output arrays out_keys[d] / out_rids[d] (HOST arrays of ndest device pointers) -- the
 * receive buffers of the destination GPUs, the local one for this rank and, for the
 * others, mappings obtained with msb64_b200_ipc_open.  The kernel's stores go over
 * NVLink / NVSwitch straight into the destination's HBM; there is no separate exchange
 * pass (the NCCL all-to-all of the unfused path).  d_cursors[d] = first slot of THIS
 * source inside destination d's arrays (the pairs lower-ranked sources send to d).  The
 * caller orders the kernel against the peers' use of their buffers (a collective before
 * and after; see distributed.py).
 *
 * msb64_b200_ipc_export / _open / _close: CUDA IPC handle (64 bytes) of an allocation made
 * with msb64_b200_device_alloc, to be sent to the other processes of the box; _open maps
 * a peer's allocation and enables peer access; returns NULL on failure. */
#define MSB64_IPC_HANDLE_BYTES 64
int msb64_b200_route_peer(const uint64_t *d_keys, const uint64_t *d_rids, uint64_t n,
			  int shift, int bits, uint64_t origin, const uint8_t *d_bin_to_dest, int ndest,
			  uint32_t *d_cursors, uint64_t *const *out_keys,
			  uint64_t *const *out_rids, void *stream);
int msb64_b200_ipc_export(void *d_ptr, void *handle64);
void *msb64_b200_ipc_open(const void *handle64);
int msb64_b200_ipc_close(void *mapped);

/* ------------------------------------- 5. the sort sharded over several GPUs
 *
 * One msb64_b200_shard per GPU: one process per GPU (the handles travel between the processes
 * as CUDA IPC handles; inplacemsdradixsort_b200/distributed.py drives this under
 * torch.distributed) or all of them in one process -- which is what sort() above does when it
 * is given numa >= 2 arrays and at least as many devices are visible: node n's pairs are sorted
 * on GPU n and node n gets the n-th key range back, the reference's contract across NUMA nodes
 * (msb_64.c:2261-2275, 1596-1606, 2180).  MSB64_B200_SINGLE_DEVICE=1 keeps sort() on one GPU,
 * MSB64_B200_VIRTUAL_SHARDS=1 lets the nodes share the visible devices (node n on device
 * n mod count).
 *
 * A step (details in csrc/msb64_shard.cuh):
 *   msb64_b200_shard_histogram   counts of the keys' top 12 bits + smallest / largest key into
 *                                the shard's histogram row (msb64_b200_shard_hist: device
 *                                pointer, msb64_b200_shard_slots() 64-bit words);
 *   [the caller all-gathers the rows of all ranks and brings them to the host]
 *   msb64_b200_shard_plan        host only: the same cut of the bin axis on every rank into
 *                                world x msb64_b200_shard_subs(world) buckets of near-equal
 *                                count; recv_caps[r] = pairs rank r can take.  Returns MSB64_OK,
 *                                MSB64_ERR_CAPACITY (a rank would overflow; msb_64.c:1574-1578)
 *                                or 1: the keys share a long prefix, the shard's digit was moved
 *                                onto their real span -- histogram again (reset = 0), gather, plan;
 *   msb64_b200_shard_exchange_sort   enqueues, without synchronising: the bucket pass (which stores
 *                                the first quarter of every peer's sub-ranges straight into the
 *                                peer's receive buffer over NVLink and stages the rest locally),
 *                                the bucket-by-bucket copies of the staged part into the peers'
 *                                receive buffers with their completion flags, and the sub-range by
 *                                sub-range sorts of what arrives (NVLink and HBM work overlap).
 * Afterwards msb64_b200_shard_keys / _rids hold msb64_b200_shard_count pairs: this rank's key
 * range, ascending, every key <= every key of the next rank.  The arrays stay valid until
 * the next step.  Between two steps every rank must have passed a collective that follows
 * its last use of the arrays (the histogram all-gather is one).
 * msb64_b200_shard_plan_host is the plan without a shard (no device needed). */
typedef struct msb64_b200_shard msb64_b200_shard;
#define MSB64_SHARD_HANDLE_BYTES 192	/* three CUDA IPC handles: receive keys, receive rids, flags */

msb64_b200_shard *msb64_b200_shard_create(int rank, int world, uint64_t capacity, double fudge);
void msb64_b200_shard_destroy(msb64_b200_shard *shard);
int msb64_b200_shard_export(msb64_b200_shard *shard, void *handles);
int msb64_b200_shard_connect_ipc(msb64_b200_shard *shard, const void *all_handles /* [world][192] */);
int msb64_b200_shard_connect_local(msb64_b200_shard *const *shards, int world);
int msb64_b200_shard_slots(void);
int msb64_b200_shard_subs(int world);
int msb64_b200_shard_histogram(msb64_b200_shard *shard, const uint64_t *d_keys, uint64_t n,
			       int reset, void *stream);
uint64_t *msb64_b200_shard_hist(msb64_b200_shard *shard);
int msb64_b200_shard_plan(msb64_b200_shard *shard, const uint64_t *all_hists /* host [world][slots] */,
			  const uint64_t *recv_caps /* [world] */, int may_retry);
int msb64_b200_shard_plan_host(const uint64_t *all_hists, int world, const uint64_t *recv_caps,
			       int may_retry, int *shift, int *bits, uint64_t *origin,
			       uint8_t *table /* [2^bits] bin -> bucket */,
			       uint64_t *counts /* [world][world * subs] */);
int msb64_b200_shard_exchange_sort(msb64_b200_shard *shard, const uint64_t *d_keys,
				   const uint64_t *d_rids, uint64_t n, void *stream, int timed);
uint64_t msb64_b200_shard_count(const msb64_b200_shard *shard);
uint64_t msb64_b200_shard_sent(const msb64_b200_shard *shard);	/* pairs the last step sent to other GPUs */
uint64_t msb64_b200_shard_sent_direct(const msb64_b200_shard *shard);	/* ... of which the route kernel stored
									   straight into the peers' memory */
uint64_t msb64_b200_shard_recv_capacity(const msb64_b200_shard *shard);
uint64_t *msb64_b200_shard_keys(msb64_b200_shard *shard);
uint64_t *msb64_b200_shard_rids(msb64_b200_shard *shard);
int msb64_b200_shard_key_range(const msb64_b200_shard *shard, uint64_t *key_lo, uint64_t *key_hi);
/* ms[0..5): route pass, wait for the first sub-range, sorting, exchange, whole step (device
 * times of the last step run with timed = 1; synchronise the stream first). */
int msb64_b200_shard_times(msb64_b200_shard *shard, double *ms);

#ifdef __cplusplus
}
#endif
#endif
