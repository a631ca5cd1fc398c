// Build shim (see numa.h).
#pragma once
#define MPOL_LOCAL 4
static inline long set_mempolicy(int mode, const unsigned long* nmask, unsigned long maxnode) {
  (void)mode; (void)nmask; (void)maxnode; return 0;
}
