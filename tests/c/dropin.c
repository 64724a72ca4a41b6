/* dropin.c -- a program written against the REFERENCE's own header, linked against
 * libmsb64_b200.so instead of msb_64.c: the drop-in boundary at the C level.
 *
 * Built by tests/c/build_dropin.py with -I/root/reference/include, i.e. it includes the
 * reference's include/msb_64.h verbatim (which expects <stdint.h> / <stddef.h> from its
 * includer, like the reference's own sources).  Only sort() and mamalloc() are used --
 * exactly what that header declares.
 *
 *   dropin <numa> <pairs per node> <fudge> <kind>     kind: 0 uniform, 1 low 24 bits, 2 few distinct
 * exit status 0 = sorted, sizes add up, checksum kept, every (key, rid) pair still together.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "msb_64.h"

static uint64_t mix(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}

int main(int argc, char **argv)
{
	const int numa = argc > 1 ? atoi(argv[1]) : 1;
	const uint64_t per = argc > 2 ? strtoull(argv[2], NULL, 0) : 100003;
	const double fudge = argc > 3 ? atof(argv[3]) : 1.5;
	const int kind = argc > 4 ? atoi(argv[4]) : 0;
	uint64_t *keys[64], *rids[64], size[64];
	uint64_t sum = 0, total = 0;
	if (numa < 1 || numa > 64) return 2;
	for (int n = 0; n < numa; ++n) {
		size[n] = per + 17 * (uint64_t) n;
		const uint64_t cap = (uint64_t) ((double) size[n] * fudge) + 8;
		keys[n] = mamalloc(cap * sizeof(uint64_t));
		rids[n] = mamalloc(cap * sizeof(uint64_t));
		if (!keys[n] || !rids[n]) return 2;
		for (uint64_t i = 0; i < size[n]; ++i) {
			uint64_t k = mix(((uint64_t) n << 40) + i + 1);
			if (kind == 1) k &= 0xFFFFFF;
			if (kind == 2) k = mix(k % 1000);
			keys[n][i] = k;
			rids[n][i] = mix(k) + 1;              /* the pair travels together: rid is a function of the key */
			sum += k;
		}
		total += size[n];
	}
	char *description[16];
	uint64_t times[16];
	sort(keys, rids, size, 64, numa, fudge, description, times);

	uint64_t after = 0, got = 0, last = 0;
	int have_last = 0;
	for (int n = 0; n < numa; ++n) {
		for (uint64_t i = 0; i < size[n]; ++i) {
			const uint64_t k = keys[n][i];
			if (have_last && k < last) {
				fprintf(stderr, "node %d element %llu: descent\n", n, (unsigned long long) i);
				return 1;
			}
			if (rids[n][i] != mix(k) + 1) {
				fprintf(stderr, "node %d element %llu: rid lost its key\n", n, (unsigned long long) i);
				return 1;
			}
			last = k;
			have_last = 1;
			after += k;
		}
		got += size[n];
	}
	if (got != total || after != sum) {
		fprintf(stderr, "pairs %llu -> %llu, checksum %llx -> %llx\n", (unsigned long long) total,
			(unsigned long long) got, (unsigned long long) sum, (unsigned long long) after);
		return 1;
	}
	for (int p = 0; p < 16 && description[p]; ++p)
		printf("%s%llu us\n", description[p], (unsigned long long) times[p]);
	printf("dropin ok: numa %d, %llu pairs\n", numa, (unsigned long long) total);
	for (int n = 0; n < numa; ++n) {
		free(keys[n]);
		free(rids[n]);
	}
	return 0;
}
