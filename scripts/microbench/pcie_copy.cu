// Microbenchmark: device -> host copy of a 25 MB grid into pageable, registered (cudaHostRegister) and
// cudaHostAlloc'ed memory, and what cudaMalloc/cudaFree of a 4 MB scratch cost per call.
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
	size_t bytes = (size_t) 1027 * 1027 * 12 * 2;
	void* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
	char* pageable = (char*) calloc(bytes, 1);
	char* reg = (char*) calloc(bytes, 1);
	double t0 = now();
	cudaError_t e = cudaHostRegister(reg, bytes, cudaHostRegisterPortable);
	printf("cudaHostRegister: %s, %.2f ms\n", cudaGetErrorString(e), (now() - t0) * 1e3);
	char* pinned; cudaHostAlloc((void**) &pinned, bytes, cudaHostAllocPortable);
	cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
	for (int rep = 0; rep < 3; rep++) {
		const char* names[3] = {"pageable", "registered", "hostalloc"};
		char* dst[3] = {pageable, reg, pinned};
		for (int k = 0; k < 3; k++) {
			cudaStreamSynchronize(s); t0 = now();
			cudaMemcpyAsync(dst[k], d, bytes, cudaMemcpyDeviceToHost, s);
			cudaStreamSynchronize(s);
			double dt = now() - t0;
			printf("D2H %-10s %.2f ms  %.1f GB/s\n", names[k], dt * 1e3, bytes / dt / 1e9);
		}
	}
	t0 = now();
	for (int k = 0; k < 10; k++) { void* p; cudaMalloc(&p, 4 << 20); cudaFree(p); }
	printf("cudaMalloc+cudaFree 4 MB: %.3f ms per pair\n", (now() - t0) * 100);
	return 0;
}
