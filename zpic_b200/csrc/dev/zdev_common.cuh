// zpic-b200 :: device runtime shared by all kernels (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "zpic_dev.h"

#define ZDEV_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "(*error*) zpic-b200 CUDA failure %s at %s:%d : %s\n", #call, __FILE__, __LINE__, \
	cudaGetErrorString(e_)); exit(-1); } } while (0)

// launch bookkeeping: every kernel launch goes through ZDEV_LAUNCH so that
// zdev_launch_count() is exact and launch errors are caught where they happen
extern uint64_t zdev_n_launch;
extern cudaStream_t zdev_strm;
extern int zdev_num_sm;
extern int zdev_time_push;

#define ZDEV_LAUNCH(kernel, grid, block, smem, ...) do { \
	kernel<<<(grid), (block), (smem), zdev_strm>>>(__VA_ARGS__); \
	zdev_n_launch++; ZDEV_CHECK(cudaGetLastError()); } while (0)

void zdev_require_init();

// 12-byte grid element, layout-compatible with the host float3 of include/em2d/zpic.h
struct f3 { float x, y, z; };

static inline int zdev_div_up(long a, long b) { return (int)((a + b - 1) / b); }
