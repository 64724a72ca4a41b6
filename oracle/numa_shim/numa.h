/* Single-node stand-in for <numa.h> (libnuma is not in this image).
 *
 * TEST INFRASTRUCTURE ONLY.  This header exists so that the UNMODIFIED reference
 * sources under /root/reference/src can be compiled into oracle/_ref/ by
 * oracle/Makefile; it is never part of the product library.  It models a machine
 * with exactly one NUMA node: every CPU belongs to node 0, binding calls are
 * no-ops, interleaved allocation is an anonymous mapping.
 *
 * Only the seven libnuma entry points msb_64.c calls are provided
 * (msb_64.c:104-108, 208, 223, 2323-2324, 2374-2375, 2424-2425).
 */
#ifndef ORACLE_NUMA_SHIM_H
#define ORACLE_NUMA_SHIM_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

struct bitmask { unsigned long size; unsigned long *maskp; };

static inline int numa_max_node(void) { return 0; }
static inline int numa_node_of_cpu(int cpu) { (void) cpu; return 0; }
static inline struct bitmask *numa_parse_nodestring(const char *s) { (void) s; return NULL; }
static inline void numa_set_membind(struct bitmask *m) { (void) m; }
static inline void numa_free_nodemask(struct bitmask *m) { (void) m; }

static inline void *numa_alloc_interleaved(size_t size)
{
	void *p = mmap(NULL, size ? size : 1, PROT_READ | PROT_WRITE,
		       MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	return p == MAP_FAILED ? NULL : p;
}

/* msb_64.c releases the sample twice (lines 2374 and 2424, the first time with
 * a 32-bit element size).  Real libnuma munmap()s, where the second call is a
 * harmless EINVAL as long as nothing was mapped there in between; to stay safe
 * with 64 live threads the shim simply leaks the (<= 4 MB) sample. */
static inline void numa_free(void *p, size_t size) { (void) p; (void) size; }

/* Reference defect the recipe has to survive: msb_64.c:2168 scans
 * `for (; d->numa_dest[p] == numa_node ; ++p);` with no bound.  For the LAST node
 * it runs off the 128-entry numa_dest array and stops only when the heap bytes
 * behind it differ from the node id; with numa == 1 (node id 0) and a fresh,
 * zeroed heap it walks on, p_to exceeds 128 and inject() (msb_64.c:1278) reads
 * sizes[]/half_block_*[] out of bounds and crashes.  msb_64.c includes <numa.h>
 * after <stdlib.h>, so the shim can give every malloc() in that translation unit
 * a 64-byte 0xFF tail without touching the source: the scan then always ends at
 * p == 128, which is what the code means. */
static inline void *oracle_shim_malloc(size_t n)
{
	char *p = (char *) malloc(n + 64);
	if (p) memset(p + n, 0xFF, 64);
	return p;
}
#define malloc(n) oracle_shim_malloc(n)

#endif
