// Microbenchmark: issue cost of packed fp32 (FMUL2/FADD2/FFMA2, sm_100) against scalar FMUL/FADD,
// alone and mixed with integer ALU work, with every operand in (opaque) registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o fp32x2_rate fp32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float2* __restrict__ in, float2 nz, int iters) {
	float2 a[ILP], b[ILP];
	int c[ILP];
	for (int i = 0; i < ILP; i++) { a[i] = in[threadIdx.x + 256 * i]; b[i] = in[threadIdx.x + 256 * (i + ILP)]; c[i] = __float_as_int(a[i].x) + i; }
	float2 z = in[threadIdx.x + 256 * 2 * ILP];
	z.x += nz.x; z.y += nz.y;          // opaque (-0,-0)
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int i = 0; i < ILP; i++) {
			if (MODE == 0) {            // scalar: 2 FMUL + 2 FADD per chain step
				a[i].x = a[i].x * b[i].x; a[i].y = a[i].y * b[i].y;
				a[i].x = a[i].x + b[i].y; a[i].y = a[i].y + b[i].x;
			} else if (MODE == 1) {     // packed: FFMA2 (exact mul, reg nz) + FADD2
				a[i] = __ffma2_rn(a[i], b[i], z);
				a[i] = __fadd2_rn(a[i], b[i]);
			} else if (MODE == 2) {     // packed, broadcast scalar operand
				a[i] = __ffma2_rn(a[i], make_float2(b[i].x, b[i].x), z);
				a[i] = __fadd2_rn(a[i], make_float2(b[i].y, b[i].y));
			} else if (MODE == 3) {     // packed, swapped halves
				a[i] = __ffma2_rn(a[i], make_float2(b[i].y, b[i].x), z);
				a[i] = __fadd2_rn(a[i], make_float2(b[i].y, b[i].x));
			} else if (MODE == 4) {     // plain FMUL2 only
				a[i] = __fmul2_rn(a[i], b[i]);
				a[i] = __fmul2_rn(a[i], b[(i + 1) % ILP]);
			} else if (MODE == 5) {     // plain FADD2 only
				a[i] = __fadd2_rn(a[i], b[i]);
				a[i] = __fadd2_rn(a[i], b[(i + 1) % ILP]);
			} else if (MODE == 6) {     // scalar + integer mix: 4 FP + 4 INT
				a[i].x = a[i].x * b[i].x; a[i].y = a[i].y * b[i].y;
				a[i].x = a[i].x + b[i].y; a[i].y = a[i].y + b[i].x;
				c[i] = (c[i] ^ (c[i] >> 3)) + 0x9e37; c[i] = (c[i] & 0xffff) | (c[i] << 5);
			} else if (MODE == 7) {     // packed + integer mix
				a[i] = __ffma2_rn(a[i], b[i], z);
				a[i] = __fadd2_rn(a[i], b[i]);
				c[i] = (c[i] ^ (c[i] >> 3)) + 0x9e37; c[i] = (c[i] & 0xffff) | (c[i] << 5);
			} else if (MODE == 8) {     // FFMA2 with the addend in the accumulator (2 reg-pair sources + self)
				a[i] = __ffma2_rn(b[i], b[(i + 1) % ILP], a[i]);
				a[i] = __ffma2_rn(b[i], b[(i + 3) % ILP], a[i]);
			}
		}
	}
	float s = 0; int t = 0;
	for (int i = 0; i < ILP; i++) { s += a[i].x + a[i].y; t += c[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int MODE>
static void run(const char* name, double inst_per_step, double lane_flop_per_step) {
	int dev = 0, sms = 0, khz = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
	float2* in; cudaMalloc(&in, sizeof(float2) * 256 * (2 * ILP + 1));
	float2 h[256 * (2 * ILP + 1)];
	for (int i = 0; i < 256 * (2 * ILP + 1); i++) h[i] = make_float2(1.0f + 1e-7f * (i % 7), 1e-9f * (i % 5));
	for (int i = 256 * 2 * ILP; i < 256 * (2 * ILP + 1); i++) h[i] = make_float2(0.0f, 0.0f);
	cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
	const int iters = 20000;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int blocks_per_sm : {2, 4}) {
		k<MODE><<<sms * blocks_per_sm, 256>>>(out, in, make_float2(-0.0f, -0.0f), 100);
		cudaEventRecord(e0);
		k<MODE><<<sms * blocks_per_sm, 256>>>(out, in, make_float2(-0.0f, -0.0f), iters);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double warps = (double) sms * blocks_per_sm * 8;
		double cyc = ms * 1e-3 * khz * 1e3;
		printf("%-40s %2d warps/SM: %7.3f ms  %.2f warp-inst/cyc/SM  %.1f fp32 lane-ops/cyc/SM\n", name, blocks_per_sm * 8, ms,
		       warps * iters * ILP * inst_per_step / cyc / sms, warps * iters * ILP * lane_flop_per_step * 32 / cyc / sms);
	}
	cudaFree(out); cudaFree(in);
}

int main() {
	run<0>("scalar 2 FMUL + 2 FADD", 4, 4);
	run<1>("FFMA2(a,b,-0) + FADD2, reg pairs", 2, 4);
	run<2>("same, scalar-broadcast operand", 2, 4);
	run<3>("same, swapped-halves operand", 2, 4);
	run<4>("FMUL2 + FMUL2", 2, 4);
	run<5>("FADD2 + FADD2", 2, 4);
	run<8>("FFMA2 + FFMA2 (accumulate)", 2, 4);
	run<6>("scalar 4 FP + 6 INT", 10, 4);
	run<7>("packed 2 FP2 + 6 INT", 8, 4);
	return 0;
}
