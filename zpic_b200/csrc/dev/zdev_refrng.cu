// zpic-b200 :: the reference's random stream, generated on the device.
//
// The reference draws every initial momentum from ONE sequential stream (em2d/random.c:48-101): two 16-bit
// multiply-with-carry generators glued into a 32-bit word, and polar Box-Muller in double precision that rejects
// the pairs outside the unit circle and caches the second deviate.  Both pieces parallelise exactly:
//   * a 16-bit multiply-with-carry generator z' = a (z & 0xffff) + (z >> 16) is the linear congruential generator
//     z' = a z mod (a 2^16 - 1) as long as 0 < z < a 2^16 - 1 (a 2^16 = 1 modulo that number, so a is the inverse of
//     the base; the excluded values are the two fixed points the reference's header warns about) - the state after
//     k draws is z a^k, one modular power;
//   * whether candidate pair t is accepted depends on that pair alone, and the r-th accepted pair yields the
//     deviates 2r and 2r+1 - a prefix sum over the acceptance flags.
// One pass counts the accepted pairs per thread (LP consecutive candidates each), a scan turns the counts into
// ranks, a second pass regenerates the candidates and writes the deviates (scaled and narrowed to float the way
// spec_set_u does: (float) (uth * rand_norm()), particles.c:97-101) to their place in the stream order.
// The double-precision arithmetic is IEEE in both builds except log(): libm's and CUDA's may differ in the last
// place of the double, which survives the narrowing to float about once in 10^9 values.
#include "zdev_common.cuh"
#include <vector>

namespace {

constexpr uint32_t AZ = 36969u, AW = 18000u;
constexpr uint32_t PZ = AZ * 65536u - 1u, PW = AW * 65536u - 1u;
constexpr int LP = 32;            // candidate pairs per thread
constexpr int RB = 256;           // threads per block

__host__ __device__ inline uint32_t mulmod(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t) (((uint64_t) a * b) % p); }
__host__ __device__ inline uint32_t powmod(uint32_t a, uint64_t e, uint32_t p) {
	uint32_t r = 1u;
	while (e) { if (e & 1u) r = mulmod(r, a, p); a = mulmod(a, a, p); e >>= 1; }
	return r;
}
__host__ __device__ inline uint32_t step_z(uint32_t z) { return AZ * (z & 0xffffu) + (z >> 16); }
__host__ __device__ inline uint32_t step_w(uint32_t w) { return AW * (w & 0xffffu) + (w >> 16); }

struct pair_out { double first, second; bool ok; };
// one candidate pair = two draws of the 32-bit word (random.c:82-90)
__device__ __forceinline__ bool candidate(uint32_t& z, uint32_t& w, double& v1, double& v2, double& rsq) {
	z = step_z(z); w = step_w(w);
	const uint32_t r1 = (z << 16) + w;
	z = step_z(z); w = step_w(w);
	const uint32_t r2 = (z << 16) + w;
	v1 = ( (double) r1 + 0.5 ) / 2147483649.0 - 1.0;
	v2 = ( (double) r2 + 0.5 ) / 2147483649.0 - 1.0;
	rsq = __dadd_rn(__dmul_rn(v1, v1), __dmul_rn(v2, v2));
	return !(rsq == 0.0 || rsq >= 1.0);
}

struct refrng_result { long long last_pair; double spare; long long accepted; };

// pass 1: accepted pairs of every thread's LP candidates -> in-block exclusive prefix + block total
__global__ void k_refrng_count(uint32_t z0, uint32_t w0, long long npairs, int* __restrict__ pre, int* __restrict__ bsum) {
	const long long tid = (long long) blockIdx.x * RB + threadIdx.x;
	const long long t0 = tid * LP;
	int c = 0;
	if (t0 < npairs) {
		uint32_t z = mulmod(z0, powmod(AZ, 2ull * (uint64_t) t0, PZ), PZ), w = mulmod(w0, powmod(AW, 2ull * (uint64_t) t0, PW), PW);
		const int n = (int) min((long long) LP, npairs - t0);
		for (int k = 0; k < n; k++) { double a, b, r; c += candidate(z, w, a, b, r) ? 1 : 0; }
	}
	__shared__ int s_w[RB / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int incl = c;
	for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	int woff = 0, tot = 0;
	for (int k = 0; k < RB / 32; k++) { const int v = s_w[k]; woff += (k < warp) ? v : 0; tot += v; }
	pre[tid] = woff + incl - c;
	if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// exclusive scan of the block totals (one block; 64-bit running sum), total -> res->accepted
__global__ void k_refrng_scan(const int* __restrict__ bsum, long long* __restrict__ boff, int nb, refrng_result* res) {
	__shared__ long long s_w[32];
	__shared__ long long s_run;
	if (threadIdx.x == 0) s_run = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
		const int i = b0 + threadIdx.x;
		const long long v = (i < nb) ? bsum[i] : 0;
		long long incl = v;
		for (int d = 1; d < 32; d <<= 1) { long long u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		long long woff = 0, tot = 0;
		for (int k = 0; k < (int) (blockDim.x >> 5); k++) { const long long c = s_w[k]; woff += (k < warp) ? c : 0; tot += c; }
		if (i < nb) boff[i] = s_run + woff + incl - v;
		__syncthreads();
		if (threadIdx.x == 0) s_run += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) res->accepted = s_run;
}

// pass 2: regenerate, write the deviates of the accepted pairs with rank < want to out[m0 + 2 rank (+1)]
// (out index m carries the scale of component (m % 3): ux, uy, uz of particle m / 3), the pair with rank want-1
// reports where the stream ends
__global__ void k_refrng_emit(uint32_t z0, uint32_t w0, long long npairs, const int* __restrict__ pre,
                              const long long* __restrict__ boff, long long want, long long m0, long long count,
                              double s0, double s1, double s2, float* __restrict__ out, refrng_result* res) {
	const long long tid = (long long) blockIdx.x * RB + threadIdx.x;
	const long long t0 = tid * LP;
	if (t0 >= npairs) return;
	long long rank = boff[blockIdx.x] + pre[tid];
	if (rank >= want) return;
	uint32_t z = mulmod(z0, powmod(AZ, 2ull * (uint64_t) t0, PZ), PZ), w = mulmod(w0, powmod(AW, 2ull * (uint64_t) t0, PW), PW);
	const int n = (int) min((long long) LP, npairs - t0);
	for (int k = 0; k < n && rank < want; k++) {
		double v1, v2, rsq;
		if (!candidate(z, w, v1, v2, rsq)) continue;
		const double fac = sqrt( __ddiv_rn(__dmul_rn(-2.0, log(rsq)), rsq) );
		const double first = __dmul_rn(v2, fac), second = __dmul_rn(v1, fac);
		const long long m = m0 + 2 * rank;
		if (out) {
			const int c = (int) (m % 3);
			out[m] = (float) __dmul_rn(c == 0 ? s0 : (c == 1 ? s1 : s2), first);
			if (m + 1 < count) { const int c1 = (c + 1) % 3; out[m + 1] = (float) __dmul_rn(c1 == 0 ? s0 : (c1 == 1 ? s1 : s2), second); }
		}
		if (rank == want - 1) { res->last_pair = t0 + k; res->spare = second; }
		rank++;
	}
}

}  // namespace

// Advance the reference stream by `count` normal deviates starting from state (z, w, have_spare, spare) - the
// variables of random.c:16-17, 69-70 - and, if d_out is given, write deviate m scaled by scale[m % 3] and narrowed
// to float to d_out[m] (device memory, count floats).  The state is updated to what the reference's would be after
// the same `count` calls of rand_norm().  Returns 0, or 1 when the state is outside the generators' linear range
// (seeds at or above a 2^16 - 1; the caller then uses the host generator).
extern "C" int zdev_ref_normals(uint32_t* zp, uint32_t* wp, int* have_spare, double* spare, long long count,
                                const float scale[3], float* d_out) {
	if (count <= 0) return 0;
	uint32_t z = *zp, w = *wp;
	if (z == 0 || z >= PZ || w == 0 || w >= PW) return 1;
	zdev_require_init();
	long long m0 = 0;
	if (*have_spare) {
		if (d_out) {
			const float v = (float) ((double) scale[0] * *spare);
			ZDEV_CHECK(cudaMemcpyAsync(d_out, &v, sizeof v, cudaMemcpyHostToDevice, zdev_strm));
			ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		}
		*have_spare = 0;
		m0 = 1;
	}
	long long want = (count - m0 + 1) / 2;            // accepted pairs still needed
	if (want == 0) return 0;
	const long long round_cap = 1ll << 28;
	const long long first_round = std::min(round_cap, (long long) ((double) want * 1.2763) + 65536);    // 4/pi + slack
	const long long max_threads = (first_round + LP - 1) / LP;
	const int max_blocks = (int) ((max_threads + RB - 1) / RB);
	int *pre = nullptr, *bsum = nullptr; long long* boff = nullptr; refrng_result* res = nullptr;
	ZDEV_CHECK(cudaMalloc(&pre, (size_t) max_blocks * RB * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&bsum, (size_t) max_blocks * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&boff, (size_t) max_blocks * sizeof(long long)));
	ZDEV_CHECK(cudaMalloc(&res, sizeof(refrng_result)));
	const double s0 = scale ? scale[0] : 0, s1 = scale ? scale[1] : 0, s2 = scale ? scale[2] : 0;
	while (want > 0) {
		const long long npairs = std::min(first_round, (long long) ((double) want * 1.2763) + 65536);
		const int nb = (int) (((npairs + LP - 1) / LP + RB - 1) / RB);
		ZDEV_LAUNCH(k_refrng_count, nb, RB, 0, z, w, npairs, pre, bsum);
		ZDEV_LAUNCH(k_refrng_scan, 1, 1024, 0, bsum, boff, nb, res);
		if (d_out)
			ZDEV_LAUNCH(k_refrng_emit, nb, RB, 0, z, w, npairs, pre, boff, want, m0, count, s0, s1, s2, d_out, res);
		else       // state only: the last needed pair is all that matters
			ZDEV_LAUNCH(k_refrng_emit, nb, RB, 0, z, w, npairs, pre, boff, want, m0, count, s0, s1, s2, (float*) nullptr, res);
		refrng_result h;
		ZDEV_CHECK(cudaMemcpyAsync(&h, res, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		long long used_pairs;
		if (h.accepted >= want) {
			used_pairs = h.last_pair + 1;
			m0 += 2 * want;
			want = 0;
			if (m0 > count) { *have_spare = 1; *spare = h.spare; }
		} else {
			used_pairs = npairs;
			m0 += 2 * h.accepted;
			want -= h.accepted;
		}
		z = mulmod(z, powmod(AZ, 2ull * (uint64_t) used_pairs, PZ), PZ);
		w = mulmod(w, powmod(AW, 2ull * (uint64_t) used_pairs, PW), PW);
	}
	ZDEV_CHECK(cudaFree(pre)); ZDEV_CHECK(cudaFree(bsum)); ZDEV_CHECK(cudaFree(boff)); ZDEV_CHECK(cudaFree(res));
	*zp = z; *wp = w;
	return 0;
}

// the host half alone (no device): the generators' state after `draws` calls of rand_uint32() (CPU tests)
extern "C" int zdev_ref_jump(uint32_t* zp, uint32_t* wp, unsigned long long draws) {
	if (*zp == 0 || *zp >= PZ || *wp == 0 || *wp >= PW) return 1;
	*zp = mulmod(*zp, powmod(AZ, draws, PZ), PZ);
	*wp = mulmod(*wp, powmod(AW, draws, PW), PW);
	return 0;
}
