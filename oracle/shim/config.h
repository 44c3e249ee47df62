// Oracle tooling only: replaces the CMake-generated config.h of the reference (include/config.h.in:1-3).
// BuDDy is provided by oracle/shim/minibuddy; Cairo by the inert oracle/shim/nullcairo (no-op drawing).
#define HAVE_BUDDY_H "TRUE"
#define HAVE_CAIRO_H "NULLCAIRO"
