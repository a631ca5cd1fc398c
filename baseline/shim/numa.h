// Build shim for the reference arm: this image has no libnuma development package.
// Single-node semantics (every CPU / GPU on NUMA node 0, no binding).  Not part of the product.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#ifdef __cplusplus
extern "C" {
#endif
struct bitmask { unsigned long size; unsigned long* maskp; };
static inline int numa_available(void) { return 0; }
static inline int numa_num_possible_cpus(void) { return 4096; }
static inline struct bitmask* numa_allocate_cpumask(void) {
  struct bitmask* b = (struct bitmask*)malloc(sizeof(struct bitmask));
  b->size = 4096; b->maskp = (unsigned long*)calloc(4096 / 8, 1); return b;
}
static inline void numa_bitmask_free(struct bitmask* b) { if (b) { free(b->maskp); free(b); } }
static inline int numa_bitmask_isbitset(const struct bitmask* b, unsigned int n) {
  return n < b->size ? (int)((b->maskp[n / (8 * sizeof(unsigned long))] >> (n % (8 * sizeof(unsigned long)))) & 1) : 0;
}
static inline int numa_node_of_cpu(int cpu) { (void)cpu; return 0; }
static inline int numa_run_on_node(int node) { (void)node; return 0; }
static inline void numa_set_preferred(int node) { (void)node; }
static inline void* numa_alloc_local(size_t size) {
  void* p = mmap(0, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  return p == MAP_FAILED ? 0 : p;
}
static inline void* numa_alloc_onnode(size_t size, int node) { (void)node; return numa_alloc_local(size); }
static inline void numa_free(void* p, size_t size) { if (p) munmap(p, size); }
#ifdef __cplusplus
}
#endif
