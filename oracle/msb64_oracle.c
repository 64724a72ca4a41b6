/* msb64_oracle.c -- CPU ORACLE for the msb_64 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A scalar, single-threaded restatement of the Polychroniou-Ross in-place MSD
 * radix sort as implemented in the reference (src/msb_64.c).  It is the checker
 * the CUDA path is compared against; nothing in the product (the package under
 * inplacemsdradixsort_b200/, include/, the C-ABI library) may call, link or
 * import it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it.
 *
 * Parity is PINNED: tests/test_oracle_pin.py checks every function below against
 * the unmodified reference compiled into oracle/_ref/libmsb64_ref.so (whose
 * helper functions are exported because the reference declares none of them
 * static) and against fixtures under tests/golden/ generated from that library
 * by tests/golden/make_golden.py.  The reference ships no tests or golden vectors
 * of its own.
 *
 * What is restated, and where it lives in the reference:
 *   orc_rand64_*           rand.c:45-85        MT19937-64 generator
 *   orc_mulhi              msb_64.c:178-186    high half of a 64x64 product
 *   orc_binary_search      msb_64.c:188-204    first index with key <= delim[i]
 *   orc_insertsort         msb_64.c:126-149
 *   orc_combsort           msb_64.c:980-1005
 *   orc_histogram          msb_64.c:701-738    digit histogram (SIMD there, scalar here)
 *   orc_partition_ip       msb_64.c:740-770    in-place cycle-following permute
 *   orc_schedule_passes    msb_64.c:1334-1400  digit widths per recursion depth
 *   orc_local_radixsort    msb_64.c:1007-1035  recursive per-bucket descent
 *   orc_extract_delimiters msb_64.c:1304-1322  sample percentiles -> splitters
 *   orc_range_delimiters   msb_64.c:1546-1564  63 sampled + 64 radix splitters
 *   orc_range_histogram    msb_64.c:239-351    range index per key + counts
 *   orc_sort               msb_64.c:1477-2259 (sort_thread) + 2261-2430 (sort)
 *
 * Deliberate differences (none changes the sorted output):
 *   - partition_ip_buf (msb_64.c:785-978) is the cache-line-buffered variant of
 *     partition_ip; it produces the same buckets with a different order inside a
 *     bucket.  The oracle always uses the unbuffered permutation.
 *   - The 64-thread / NUMA block shuffling (combine, compact, balance, block swap,
 *     inject: msb_64.c:1220-1302, 1674-2198) is the parallel in-place realisation
 *     of "range-partition into 128 ranges"; the oracle does that step with a
 *     scratch copy.
 *   - The reference refuses inputs under 2^25 pairs (msb_64.c:1569) and sample
 *     sizes of 0; the oracle extends the same algorithm downward so that small
 *     and empty inputs can be checked.
 *   - The reference leaves the sampling seed uninitialised (thread_data_t.seed is
 *     never written); the oracle takes it as an argument.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ rand.c:45-85 */
typedef struct { uint64_t num[313]; size_t index; } orc_rand64_t;

void orc_rand64_seed(orc_rand64_t *s, uint64_t seed)
{
	s->num[0] = seed;
	for (size_t i = 1; i < 312; ++i) {
		uint64_t prev = s->num[i - 1];
		s->num[i] = 6364136223846793005ull * (prev ^ (prev >> 62)) + i;
	}
	s->index = 312;
}

uint64_t orc_rand64_next(orc_rand64_t *s)
{
	uint64_t *n = s->num;
	if (s->index == 312) {
		const uint64_t upper = 0xffffffff80000000ull, lower = 0x7fffffffull;
		const uint64_t twist = 0xb5026f5aa96619e9ull;
		for (size_t i = 0; i < 312; ++i) {
			/* n[312] mirrors n[0] for the wrap-around element (rand.c:69) */
			uint64_t nxt = (i + 1 == 312) ? n[0] : n[i + 1];
			uint64_t x = (n[i] & upper) | (nxt & lower);
			uint64_t far = n[(i + 156) % 312];
			n[i] = far ^ (x >> 1) ^ ((x & 1) ? twist : 0);
		}
		s->index = 0;
	}
	uint64_t x = n[s->index++];
	x ^= (x >> 29) & 0x5555555555555555ull;
	x ^= (x << 17) & 0x71d67fffeda60000ull;
	x ^= (x << 37) & 0xfff7eee000000000ull;
	x ^= (x >> 43);
	return x;
}

/* fill an array with the generator's stream (how config[0] "rand.c generator"
 * data is produced) */
void orc_rand64_fill(uint64_t seed, uint64_t *out, uint64_t count)
{
	orc_rand64_t s;
	orc_rand64_seed(&s, seed);
	for (uint64_t i = 0; i < count; ++i) out[i] = orc_rand64_next(&s);
}

/* ------------------------------------------------------------- msb_64.c:178-186 */
uint64_t orc_mulhi(uint64_t x, uint64_t y)
{
	return (uint64_t) (((unsigned __int128) x * y) >> 64);
}

/* ------------------------------------------------------------- msb_64.c:188-204 */
uint64_t orc_binary_search(const uint64_t *delim, uint64_t count, uint64_t key)
{
	uint64_t lo = 0, hi = count;
	while (lo < hi) {
		uint64_t mid = (lo + hi) >> 1;
		if (key > delim[mid]) lo = mid + 1; else hi = mid;
	}
	return lo;
}

/* ------------------------------------------------------------- msb_64.c:126-149 */
void orc_insertsort(uint64_t *keys, uint64_t *rids, uint64_t size)
{
	for (uint64_t i = 1; i < size; ++i) {
		uint64_t k = keys[i], r = rids[i];
		uint64_t j = i;
		/* strict '<': equal keys keep their order, as in the reference */
		while (j > 0 && k < keys[j - 1]) {
			keys[j] = keys[j - 1];
			rids[j] = rids[j - 1];
			--j;
		}
		keys[j] = k;
		rids[j] = r;
	}
}

/* ------------------------------------------------------------ msb_64.c:980-1005 */
void orc_combsort(uint64_t *keys, uint64_t *rids, uint64_t size)
{
	if (size < 2) return;
	const float shrink = 0.77f;
	uint64_t gap = (uint64_t) (size * shrink);
	if (gap == 0) gap = 1;	/* size >= 2 gives gap >= 1 already; guard only */
	for (;;) {
		int swapped = 0;
		for (uint64_t i = 0, j = gap; j < size; ++i, ++j) {
			if (keys[i] > keys[j]) {
				uint64_t t = keys[i]; keys[i] = keys[j]; keys[j] = t;
				t = rids[i]; rids[i] = rids[j]; rids[j] = t;
				swapped = 1;
			}
		}
		if (gap > 1) gap = (uint64_t) (gap * shrink);
		else if (!swapped) break;
	}
}

/* ------------------------------------------------------------- msb_64.c:701-738 */
void orc_histogram(const uint64_t *keys, uint64_t size, uint64_t *count,
		   uint8_t shift_bits, uint8_t radix_bits)
{
	uint64_t buckets = 1ull << radix_bits, mask = buckets - 1;
	memset(count, 0, buckets * sizeof(uint64_t));
	for (uint64_t i = 0; i < size; ++i)
		count[(keys[i] >> shift_bits) & mask]++;
}

/* ------------------------------------------------------------- msb_64.c:740-770
 * American-flag style permutation: bucket b owns [start_b, end_b); offsets[b]
 * starts at end_b and walks down as slots are filled.  A cycle starts at the
 * first unfilled slot `e` and follows displaced items until one lands on `e`. */
void orc_partition_ip(uint64_t *keys, uint64_t *rids, uint64_t size,
		      const uint64_t *sizes, uint64_t *offsets,
		      uint8_t shift_bits, uint8_t radix_bits)
{
	uint64_t buckets = 1ull << radix_bits, mask = buckets - 1;
	uint64_t run = 0, b, e = 0, slot;
	if (size == 0) return;
	for (b = 0; b < buckets; ++b) { run += sizes[b]; offsets[b] = run; }
	b = 0;
	while (sizes[b] == 0) ++b;
	do {
		uint64_t k = keys[e], r = rids[e];
		do {
			slot = --offsets[(k >> shift_bits) & mask];
			uint64_t tk = keys[slot], tr = rids[slot];
			keys[slot] = k; rids[slot] = r;
			k = tk; r = tr;
		} while (slot != e);
		/* skip buckets whose fill pointer has reached their start */
		do { e += sizes[b++]; } while (b != buckets && e == offsets[b]);
	} while (b != buckets);
}

/* ----------------------------------------------------------- msb_64.c:1324-1332 */
static uint8_t ceil_log2_u64(uint64_t x)
{
	uint8_t p = 0;
	while ((1ull << p) < x) p++;
	return p;
}
static uint64_t ceil_div_u64(uint64_t x, uint64_t y) { return (x + y - 1) / y; }

/* ----------------------------------------------------------- msb_64.c:1334-1400
 * Plans the digit width of every recursion depth so that the leaves hold at most
 * cache_limit (6500) pairs.  buffered[d]: 1 = out-of-cache (buffered) partition,
 * 0 = in-cache partition, -1 = sentinel (comb sort whatever bits are left).
 * Returns the number of partitioning depths; -1 where the reference asserts. */
int orc_schedule_passes(uint64_t size, int8_t bits, int8_t *radix_bits, int8_t *buffered)
{
	const uint64_t cache_limit = 6500;
	int p = 0;
	int8_t lp = (int8_t) ceil_log2_u64(ceil_div_u64(size, cache_limit));
	if (!(lp < bits)) return -1;
	if (size <= cache_limit) {
		/* nothing: fits already */
	} else if (lp <= 5) {			/* one in-cache split, 8..32 way */
		if (lp < 3) lp = 3;
		if (lp > bits) lp = bits;
		buffered[p] = 0; radix_bits[p++] = lp;
	} else if (lp <= 9) {			/* one buffered split, up to 512 way */
		buffered[p] = 1; radix_bits[p++] = lp;
	} else if (lp <= 12) {			/* 8-way in-cache, then buffered */
		buffered[p] = 0; radix_bits[p++] = 3;
		buffered[p] = 1; radix_bits[p++] = lp - 3;
	} else if (lp <= 14) {			/* buffered, then 32-way in-cache */
		buffered[p] = 1; radix_bits[p++] = lp - 5;
		buffered[p] = 0; radix_bits[p++] = 5;
	} else if (lp <= 18) {			/* two buffered splits */
		buffered[p] = 1; radix_bits[p++] = lp >> 1;
		buffered[p] = 1; radix_bits[p++] = (lp + 1) >> 1;
	} else if (lp <= 27) {			/* three buffered splits */
		buffered[p] = 1; radix_bits[p++] = lp / 3;
		lp -= lp / 3;
		buffered[p] = 1; radix_bits[p++] = lp >> 1;
		buffered[p] = 1; radix_bits[p++] = (lp + 1) >> 1;
	} else return -1;
	for (int i = 0; i < p; ++i) { size >>= radix_bits[i]; bits -= radix_bits[i]; }
	if (size > cache_limit) return -1;
	int last = (int) ceil_log2_u64(size) - 2;	/* final in-cache split */
	if (last > bits) last = bits;
	bits -= last;
	buffered[p] = 0; radix_bits[p++] = (int8_t) last;
	buffered[p] = -1; radix_bits[p] = bits;		/* seal */
	return p;
}

/* ----------------------------------------------------------- msb_64.c:1007-1035
 * bits[d] is the number of still-unsorted low key bits on entry to depth d
 * (the caller turns schedule_passes' widths into suffix sums, msb_64.c:2243). */
void orc_local_radixsort(uint64_t *keys, uint64_t *rids, uint64_t size,
			 const int8_t *bits, const int8_t *buffered, int depth,
			 uint64_t **hist, uint64_t **offsets)
{
	if (size <= 20) { orc_insertsort(keys, rids, size); return; }
	if (buffered[depth] < 0) { orc_combsort(keys, rids, size); return; }
	int8_t shift = bits[depth + 1];
	int8_t width = bits[depth] - shift;
	uint64_t buckets = 1ull << width;
	orc_histogram(keys, size, hist[depth], shift, width);
	orc_partition_ip(keys, rids, size, hist[depth], offsets[depth], shift, width);
	if (shift == 0) return;
	uint64_t at = 0;
	for (uint64_t b = 0; b < buckets; ++b) {
		uint64_t cnt = hist[depth][b];
		orc_local_radixsort(keys + at, rids + at, cnt, bits, buffered,
				    depth + 1, hist, offsets);
		at += cnt;
	}
}

/* sort one range whose keys agree above bit `bits` (msb_64.c:2232-2252) */
int orc_sort_range(uint64_t *keys, uint64_t *rids, uint64_t size, int8_t bits)
{
	int8_t widths[8], buffered[8];
	uint64_t *hist[5], *offs[5];
	if (size == 0) return 0;
	int depths = orc_schedule_passes(size, bits, widths, buffered);
	if (depths < 0) return -1;
	for (int i = depths; i-- > 0; ) widths[i] += widths[i + 1];
	for (int i = 0; i < 5; ++i) {
		hist[i] = malloc(4096 * sizeof(uint64_t));
		offs[i] = malloc(4096 * sizeof(uint64_t));
	}
	orc_local_radixsort(keys, rids, size, widths, buffered, 0, hist, offs);
	for (int i = 0; i < 5; ++i) { free(hist[i]); free(offs[i]); }
	return 0;
}

/* ----------------------------------------------------------- msb_64.c:1304-1322
 * delimiter[] arrives zero-filled with ~0 at the last used slot; the number of
 * splitters is the index of that ~0.  A splitter that sits in a run of equal
 * sample values is decremented when more of the run lies after it than before,
 * so that the run goes to the next range as a whole. */
void orc_extract_delimiters(const uint64_t *sample, uint64_t sample_size,
			    uint64_t *delimiter)
{
	uint64_t parts = 0;
	while (delimiter[parts] != ~(uint64_t) 0) parts++;
	double percentile = sample_size * 1.0 / (parts + 1);
	for (uint64_t i = 0; i < parts; ++i) {
		uint64_t index = (uint64_t) (percentile * (i + 1) - 0.001);
		uint64_t d = sample[index], start, end;
		for (start = index; start; --start)
			if (sample[start] != d) break;
		for (end = index; end != sample_size; ++end)
			if (sample[end] != d) break;
		if (index - start < end - index && d) d--;
		delimiter[i] = d;
	}
}

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
	return x < y ? -1 : x > y;
}

/* ----------------------------------------------------------- msb_64.c:1546-1564
 * 128 range splitters = 63 sampled thread splitters + ~0, merged with the 64
 * radix splitters (p << 58) - 1 (p = 0 wraps to ~0), sorted.  Every range thus
 * lies inside one value of the top 6 key bits, which is why the local sort can
 * start at bit 58.  thread_delim (64 entries) receives the sampled splitters. */
void orc_range_delimiters(const uint64_t *sorted_sample, uint64_t sample_size,
			  uint64_t *thread_delim, uint64_t *range_delim)
{
	memset(thread_delim, 0, 64 * sizeof(uint64_t));
	thread_delim[63] = ~(uint64_t) 0;
	orc_extract_delimiters(sorted_sample, sample_size, thread_delim);
	for (uint64_t p = 0; p < 64; ++p) {
		range_delim[p] = thread_delim[p];
		range_delim[p + 64] = (p << 58) - 1;
	}
	qsort(range_delim, 128, sizeof(uint64_t), cmp_u64);
}

/* ------------------------------------------------------------- msb_64.c:239-351
 * range of a key = first r with key <= delim[r]; the reference finds it with a
 * 3-level SIMD search over 128 splitters, which is this binary search. */
void orc_range_histogram(const uint64_t *keys, uint8_t *ranges, uint64_t size,
			 uint64_t *count, const uint64_t *delim)
{
	for (uint64_t i = 0; i < size; ++i) {
		uint64_t r = orc_binary_search(delim, 128, keys[i]);
		ranges[i] = (uint8_t) r;
		count[r]++;
	}
}

/* ------------------------------------------- msb_64.c:1477-2259 and 2261-2430
 * Whole sort for `numa` arrays.  keys[n]/rids[n] hold size[n] pairs on entry and
 * must have room for what the node receives (the reference's fudge factor); on
 * return node n holds the keys of thread ranges [n*64/numa, (n+1)*64/numa) in
 * ascending order and size[n] is updated (msb_64.c:1555-1557, 2180).
 * capacity[n] bounds what may be written.  Returns 0, or -1 when a node would
 * overflow its capacity / the pass schedule is out of the reference's range. */
int orc_sort(uint64_t **keys, uint64_t **rids, uint64_t *size, const uint64_t *capacity,
	     int numa, uint64_t seed)
{
	uint64_t total = 0;
	for (int n = 0; n < numa; ++n) total += size[n];
	if (total == 0) return 0;
	if (numa < 1 || 64 % numa) return -1;

	/* sample (msb_64.c:1515-1521, 2320-2322) */
	uint64_t sample_size = (uint64_t) (0.005 * total);
	if (sample_size > 500000) sample_size = 500000;
	if (sample_size < 256) sample_size = total < 256 ? total : 256;	/* oracle extension */
	uint64_t *sample = malloc(sample_size * sizeof(uint64_t));
	orc_rand64_t gen;
	orc_rand64_seed(&gen, seed);
	for (uint64_t i = 0; i < sample_size; ++i) {
		uint64_t at = orc_mulhi(orc_rand64_next(&gen), total);
		int n = 0;
		while (at >= size[n]) at -= size[n++];
		sample[i] = keys[n][at];
	}
	qsort(sample, sample_size, sizeof(uint64_t), cmp_u64);	/* 8 LSB passes there */

	uint64_t thread_delim[64], range_delim[128];
	orc_range_delimiters(sample, sample_size, thread_delim, range_delim);
	free(sample);

	/* range partition (scratch copy stands in for the block shuffle) */
	uint64_t *tk = malloc(total * sizeof(uint64_t));
	uint64_t *tr = malloc(total * sizeof(uint64_t));
	uint8_t *rg = malloc(total);
	uint64_t count[128] = {0}, start[129];
	uint64_t at = 0;
	for (int n = 0; n < numa; ++n) {
		memcpy(tk + at, keys[n], size[n] * sizeof(uint64_t));
		memcpy(tr + at, rids[n], size[n] * sizeof(uint64_t));
		at += size[n];
	}
	orc_range_histogram(tk, rg, total, count, range_delim);
	start[0] = 0;
	for (int r = 0; r < 128; ++r) start[r + 1] = start[r] + count[r];

	/* node of a range (msb_64.c:1596-1606): ranges up to and including the one
	 * ending at numa_delimiter[n] = thread_delim[(n+1)*64/numa - 1] go to node n */
	int node_of[128];
	{
		int r = 0, per = 64 / numa;
		for (int n = 0; n + 1 < numa; ++n) {
			uint64_t q = orc_binary_search(range_delim, 128, thread_delim[(n + 1) * per - 1]);
			for (; (uint64_t) r <= q && r < 128; ++r) node_of[r] = n;
		}
		for (; r < 128; ++r) node_of[r] = numa - 1;
	}
	int rc = 0;
	uint64_t new_size[64] = {0};
	for (int r = 0; r < 128; ++r) new_size[node_of[r]] += count[r];
	for (int n = 0; n < numa; ++n)
		if (new_size[n] > capacity[n]) rc = -1;
	if (rc == 0) {
		uint64_t fill[64] = {0}, cursor[128];
		for (int r = 0; r < 128; ++r) {
			cursor[r] = fill[node_of[r]];
			fill[node_of[r]] += count[r];
		}
		uint64_t pos[128];
		memcpy(pos, cursor, sizeof(pos));
		for (uint64_t i = 0; i < total; ++i) {
			int r = rg[i], n = node_of[r];
			keys[n][pos[r]] = tk[i];
			rids[n][pos[r]] = tr[i];
			pos[r]++;
		}
		/* local MSB radix sort of every range on 58 bits (msb_64.c:2239-2248) */
		for (int r = 0; r < 128 && rc == 0; ++r) {
			int n = node_of[r];
			rc = orc_sort_range(keys[n] + cursor[r], rids[n] + cursor[r], count[r], 58);
		}
		for (int n = 0; n < numa; ++n) size[n] = new_size[n];
	}
	free(tk); free(tr); free(rg);
	return rc;
}

/* ------------------------------------------------------------ checking helpers
 * (the reference's check(), msb_64.c:2432-2505: ascending keys + sum checksum).
 * orc_pair_digest is an order-independent digest of the (key, rid) multiset so
 * that "same pairs, any order among equal keys" can be compared at full size. */
uint64_t orc_check_sorted(const uint64_t *keys, uint64_t size, uint64_t *checksum)
{
	uint64_t sum = 0, bad = 0, prev = 0;
	for (uint64_t i = 0; i < size; ++i) {
		if (keys[i] < prev) bad++;
		prev = keys[i];
		sum += keys[i];
	}
	if (checksum) *checksum = sum;
	return bad;
}

static uint64_t mix64(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

uint64_t orc_pair_digest(const uint64_t *keys, const uint64_t *rids, uint64_t size)
{
	uint64_t acc = 0;
	for (uint64_t i = 0; i < size; ++i)
		acc += mix64(keys[i] + 0x9e3779b97f4a7c15ull * mix64(rids[i] + 1));
	return acc;
}

/* order-DEPENDENT digest of the key sequence (bit-exact key output) */
uint64_t orc_key_sequence_digest(const uint64_t *keys, uint64_t size)
{
	uint64_t acc = 0x243f6a8885a308d3ull;
	for (uint64_t i = 0; i < size; ++i)
		acc = mix64(acc ^ keys[i]) + i;
	return acc;
}
