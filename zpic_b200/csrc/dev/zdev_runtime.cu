// zpic-b200 :: device runtime (stream, events, launch accounting)
#include "zdev_common.cuh"
#include <cstring>

uint64_t zdev_n_launch = 0;
cudaStream_t zdev_strm = nullptr;
int zdev_num_sm = 148;
static int zdev_is_ready = 0;
static void* zdev_flush_buf = nullptr;
static size_t zdev_flush_bytes = 0;

void zdev_require_init() {
	if (!zdev_is_ready) {
		if (zdev_init(-1) != 0) {
			fprintf(stderr, "(*error*) zpic-b200: no usable CUDA device; this build has no CPU path, aborting.\n");
			exit(-1);
		}
	}
}

extern "C" int zdev_init(int device) {
	if (zdev_is_ready) return 0;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 1;
	if (device < 0) {
		const char* e = getenv("ZPIC_DEVICE");
		if (!e) e = getenv("LOCAL_RANK");
		device = e ? atoi(e) : 0;
	}
	device %= n;
	if (cudaSetDevice(device) != cudaSuccess) return 2;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 3;
	zdev_num_sm = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&zdev_strm, cudaStreamNonBlocking) != cudaSuccess) return 4;
	zdev_is_ready = 1;
	return 0;
}

// run on a stream owned by the caller (e.g. torch's current stream so NCCL collectives issued by
// torch.distributed are ordered with the kernels); nullptr restores the library stream
static cudaStream_t zdev_own_strm = nullptr;
extern "C" void zdev_set_stream(void* stream) {
	zdev_require_init();
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	if (!zdev_own_strm) zdev_own_strm = zdev_strm;
	zdev_strm = stream ? (cudaStream_t) stream : zdev_own_strm;
}

// per-launch CUDA-event timing of the push kernels (bench roofline), see zdev_spec*_push_timing
int zdev_time_push = 0;
extern "C" void zdev_set_push_timing(int on) { zdev_time_push = on; }

extern "C" int zdev_ready(void) { return zdev_is_ready; }
extern "C" void zdev_sync(void) { zdev_require_init(); ZDEV_CHECK(cudaStreamSynchronize(zdev_strm)); }
extern "C" void* zdev_stream(void) { zdev_require_init(); return (void*) zdev_strm; }
extern "C" uint64_t zdev_launch_count(void) { return zdev_n_launch; }

extern "C" void* zdev_event_create(void) {
	zdev_require_init();
	cudaEvent_t ev; ZDEV_CHECK(cudaEventCreate(&ev)); return (void*) ev;
}
extern "C" void zdev_event_record(void* ev) { ZDEV_CHECK(cudaEventRecord((cudaEvent_t) ev, zdev_strm)); }
extern "C" float zdev_event_elapsed_ms(void* a, void* b) {
	float ms = 0; ZDEV_CHECK(cudaEventSynchronize((cudaEvent_t) b));
	ZDEV_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t) a, (cudaEvent_t) b)); return ms;
}
extern "C" void zdev_event_destroy(void* ev) { ZDEV_CHECK(cudaEventDestroy((cudaEvent_t) ev)); }

extern "C" void zdev_mem_info(size_t* f, size_t* t) { zdev_require_init(); ZDEV_CHECK(cudaMemGetInfo(f, t)); }

extern "C" void zdev_flush_l2(void) {
	zdev_require_init();
	if (!zdev_flush_buf) {
		zdev_flush_bytes = (size_t) 256 << 20;   // 2x the 126 MB L2
		ZDEV_CHECK(cudaMalloc(&zdev_flush_buf, zdev_flush_bytes));
	}
	ZDEV_CHECK(cudaMemsetAsync(zdev_flush_buf, 0, zdev_flush_bytes, zdev_strm));
}

// Host buffers that are copied to / from the device again and again (the E, B, J mirrors of a simulation) are
// page-locked on first use: a pageable cudaMemcpy runs at ~6 GB/s on this box, a pinned one at PCIe speed.
// The owner of the buffer calls zdev_host_unpin() before freeing it.
struct zdev_pin { const void* ptr; size_t bytes; };
static zdev_pin zdev_pins[32];
extern "C" void zdev_host_pin(const void* ptr, size_t bytes) {
	if (!ptr || bytes < ((size_t) 1 << 20)) return;
	zdev_pin* slot = nullptr;
	for (auto& p : zdev_pins) { if (p.ptr == ptr) return; if (!p.ptr && !slot) slot = &p; }
	if (!slot) return;
	if (cudaHostRegister((void*) ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return; }
	slot->ptr = ptr; slot->bytes = bytes;
}
extern "C" void zdev_host_unpin(const void* ptr) {
	if (!ptr) return;
	for (auto& p : zdev_pins) if (p.ptr == ptr) { cudaHostUnregister((void*) ptr); p.ptr = nullptr; p.bytes = 0; }
}
