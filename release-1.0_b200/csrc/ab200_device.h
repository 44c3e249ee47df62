// ab200_device.h -- launcher prototypes between the C-ABI host code and ab200_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ab200_evolve.cuh"
#include "ab200_grid.cuh"


namespace ab200 { inline namespace AB_VARIANT {
int dev_upload_system(const SystemDesc* s, char* errbuf, int errlen);
// variant-independent signatures (used to reach the AB_KW=2 build from host code compiled against the AB_KW=1 headers)
int dev_upload_system_v(const void* system_desc, char* errbuf, int errlen);
int dev_evolve_capacity_v(char* errbuf, int errlen);
int dev_launch_evolve_v(cudaStream_t st, int n_cells, int slots, const int32_t* cell_loc, const double* cell_box, const void* evolve_params,
                        uint64_t* arena, int arena_words, const uint32_t* binom, int* next_cell,
                        double* reach, int32_t* reach_loc, int32_t* n_reach,
                        double* fin, int32_t* fin_loc, int32_t* n_fin, int32_t* status, int32_t* steps, int32_t* peak, double* flops,
                        const uint64_t* enc_in, const uint64_t* enc_in_off, uint64_t* enc_out, int enc_cap, int32_t* enc_used,
                        char* errbuf, int errlen);
int dev_set_stack(size_t bytes, char* errbuf, int errlen);
int dev_launch_flow_bounds(cudaStream_t st, int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                           double* h_out, double* b_out, int32_t* status, char* errbuf, int errlen);
int dev_launch_integrator_flow_bounds(cudaStream_t st, int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                                      double* h_out, double* b_out, int32_t* status, char* errbuf, int errlen);
int dev_launch_integrator_flow(cudaStream_t st, int loc, const double* box, double hmax, int order, uint64_t* arena, int arena_words,
                               const uint32_t* binom, uint64_t* out, int out_words, char* errbuf, int errlen);
int dev_launch_jump_check(cudaStream_t st, int n, int blocks, int loc, int ev, const double* boxes, uint64_t* arena, int arena_words,
                          const uint32_t* binom, int* next, int32_t* tri, int32_t* status, int32_t* lie_missing, char* errbuf, int errlen);
int dev_launch_tm_op(cudaStream_t st, const uint64_t* in, uint64_t* arena, int arena_words, const uint32_t* binom, uint64_t* out, int out_words,
                     char* errbuf, int errlen);
int dev_launch_interval_op(cudaStream_t st, int n, const double* in, double* out, char* errbuf, int errlen);
int dev_launch_fill(cudaStream_t st, uint64_t* p, size_t n, uint64_t v, char* errbuf, int errlen);
int dev_launch_raster(cudaStream_t st, int n_cells, int max_boxes, const double* boxes, const int32_t* box_loc, const int32_t* n_boxes,
                      const KeyLayout& K, int n_locs, uint64_t* table, uint64_t mask, unsigned long long* counters, char* errbuf, int errlen);
int dev_launch_compact(cudaStream_t st, const uint64_t* table, size_t n, uint64_t* out, unsigned long long* count, char* errbuf, int errlen);
int dev_radix_sort(cudaStream_t st, uint64_t* keys, uint64_t* tmp, size_t n, int bits, unsigned long long varying, unsigned* hist, int* result_in_tmp,
                   char* errbuf, int errlen);
int dev_sort_tile();
int dev_unique_sorted(cudaStream_t st, const uint64_t* in, size_t n, unsigned* pos, uint64_t* out, char* errbuf, int errlen);
int dev_fp64_peak(cudaStream_t st, double* scratch, int blocks, int threads, int iters, int fma, char* errbuf, int errlen);
} }
