// context.cu -- process-wide device context, memory helpers, timers.
#include <cstdarg>
#include <cstring>
#include <string>

#include "common.cuh"

namespace nbgpu {

static thread_local std::string g_error;
// One context per process by default.  A host thread that calls nbgpu_thread_bind_device(d) switches
// to a context of its own for device d (the single-process multi-GPU mode: one worker thread per GPU,
// each running the unchanged one-GPU-per-rank code); every other thread keeps the process context.
constexpr int kMaxDeviceContexts = 16;
static Context g_default_ctx;
static Context g_device_ctx[kMaxDeviceContexts];
static thread_local Context *t_ctx = &g_default_ctx;
#define g_ctx (*t_ctx)

Context &ctx() { return *t_ctx; }

cudaError_t dmalloc_bytes(void **p, size_t bytes)
{
	Context &c = g_ctx;
	if (!bytes)
		bytes = 1;
	if (!c.pool)
		return cudaMalloc(p, bytes);
	cudaError_t e = cudaMallocAsync(p, bytes, c.stream);
	if (e == cudaErrorMemoryAllocation) {
		// blocks parked in the pool may be too fragmented for this request: give them back, retry
		cudaGetLastError();
		cudaStreamSynchronize(c.stream);
		cudaMemPoolTrimTo(c.pool, 0);
		e = cudaMallocAsync(p, bytes, c.stream);
	}
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	return e;
}

cudaError_t dfree(void *p)
{
	Context &c = g_ctx;
	if (!p)
		return cudaSuccess;
	if (!c.pool)
		return cudaFree(p);
	return cudaFreeAsync(p, c.stream);
}

void set_error(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_error = buf;
}

static int init_device(int device)
{
	Context &c = *t_ctx;
	if (c.ready)
		return NBGPU_OK;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		// no CPU fallback: the product path needs a CUDA device
		set_error("no usable CUDA device (%s); libnbgpu has no CPU fallback",
			  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	if (device < 0) {
		const char *lr = getenv("LOCAL_RANK");
		device = lr ? atoi(lr) % count : 0;
	}
	NB_ARG(device < count);
	NB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	NB_CUDA(cudaGetDeviceProperties(&prop, device));
	c.device = device;
	c.sm_count = prop.multiProcessorCount;
	NB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
	NB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
	NB_CUDA(cudaEventCreate(&c.ev_a));
	NB_CUDA(cudaEventCreate(&c.ev_b));
	NB_CUDA(cudaEventCreateWithFlags(&c.ev_stage[0], cudaEventDisableTiming));
	NB_CUDA(cudaEventCreateWithFlags(&c.ev_stage[1], cudaEventDisableTiming));
	c.pool = nullptr;
	if (!getenv("NBGPU_NO_POOL")) {
		int supported = 0;
		cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device);
		if (supported && cudaDeviceGetDefaultMemPool(&c.pool, device) == cudaSuccess) {
			uint64_t keep = UINT64_MAX;
			if (cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
				c.pool = nullptr;
		} else {
			c.pool = nullptr;
		}
		cudaGetLastError();
	}
	NB_CUDA(nbgpu::dmalloc(&c.partials, sizeof(double) * 2 * kPartialRegion));
	c.ready = true;
	return NBGPU_OK;
}

int ensure_init()
{
	if (g_ctx.ready) {
		// ctypes / host threads other than the initialising one
		cudaSetDevice(g_ctx.device);
		return NBGPU_OK;
	}
	return init_device(-1);
}

int ensure_stage(size_t bytes)
{
	Context &c = g_ctx;
	if (c.stage_bytes >= bytes)
		return NBGPU_OK;
	NB_CUDA(cudaStreamSynchronize(c.copy_stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	for (int i = 0; i < 2; i++) {
		if (c.stage[i])
			NB_CUDA(cudaFreeHost(c.stage[i]));
		c.stage[i] = nullptr;
	}
	c.stage_bytes = 0;
	for (int i = 0; i < 2; i++)
		NB_CUDA(cudaMallocHost(&c.stage[i], bytes));
	c.stage_bytes = bytes;
	return NBGPU_OK;
}

int ensure_workspace(size_t bytes)
{
	Context &c = g_ctx;
	if (c.ws_bytes >= bytes)
		return NBGPU_OK;
	NB_CUDA(cudaStreamSynchronize(c.stream));
	if (c.ws)
		NB_CUDA(nbgpu::dfree(c.ws));
	c.ws = nullptr;
	c.ws_bytes = 0;
	cudaError_t e = nbgpu::dmalloc(&c.ws, bytes);
	if (e != cudaSuccess) {
		set_error("workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
		cudaGetLastError();
		return NBGPU_ERR_NOMEM;
	}
	c.ws_bytes = bytes;
	return NBGPU_OK;
}

// Pin [ptr, ptr+bytes) in the persisting-L2 carve-out for kernels on the stream; everything else is
// marked streaming (`others_streaming`) or left to the normal policy.  Returns whether a window was
// installed.
bool l2_pin(void *ptr, size_t bytes, bool others_streaming)
{
	if (getenv("NBGPU_NO_L2_PIN"))
		return false;
	Context &c = ctx();
	int max_persist = 0, max_window = 0;
	cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
	cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
	if (max_persist <= 0 || max_window <= 0 || bytes == 0 || bytes > (size_t)max_window ||
	    bytes > (size_t)max_persist)
		return false;   // vectors larger than the carve-out: leave the L2 to its own policy
	if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	v.accessPolicyWindow.base_ptr = ptr;
	v.accessPolicyWindow.num_bytes = bytes;
	v.accessPolicyWindow.hitRatio = 1.0f;
	v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	v.accessPolicyWindow.missProp = others_streaming ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
	if (cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return true;
}

void l2_unpin()
{
	Context &c = ctx();
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	v.accessPolicyWindow.num_bytes = 0;
	cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &v);
	cudaCtxResetPersistingL2Cache();
	// give the carve-out back: left in place it takes 79 of the 126 MB of L2 away from every later
	// kernel of the process (measured: a 4 M-dof solve after a 1 M-dof one ran at 280 instead of
	// 164 us per iteration)
	cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
	cudaGetLastError();
}

}  // namespace nbgpu

using namespace nbgpu;

extern "C" {

int nbgpu_init(int device)
{
	if (g_ctx.ready) {
		if (device >= 0 && device != g_ctx.device) {
			set_error("already bound to device %d", g_ctx.device);
			return NBGPU_ERR_ARG;
		}
		return NBGPU_OK;
	}
	return init_device(device);
}

int nbgpu_finalize(void)
{
	Context &c = g_ctx;
	if (!c.ready)
		return NBGPU_OK;
	cudaSetDevice(c.device);
	cudaStreamSynchronize(c.stream);
	cudaStreamSynchronize(c.copy_stream);
	for (int i = 0; i < 2; i++) {
		if (c.stage[i])
			cudaFreeHost(c.stage[i]);
		cudaEventDestroy(c.ev_stage[i]);
		if (c.poll_ev[i])
			cudaEventDestroy(c.poll_ev[i]);
	}
	if (c.ws)
		nbgpu::dfree(c.ws);
	if (c.partials)
		nbgpu::dfree(c.partials);
	if (c.dev_state)
		nbgpu::dfree(c.dev_state);
	if (c.host_state)
		cudaFreeHost(c.host_state);
	if (c.pool) {
		cudaStreamSynchronize(c.stream);
		cudaMemPoolTrimTo(c.pool, 0);
		c.pool = nullptr;
	}
	cudaEventDestroy(c.ev_a);
	cudaEventDestroy(c.ev_b);
	cudaStreamDestroy(c.stream);
	cudaStreamDestroy(c.copy_stream);
	uint64_t launches = c.launches;
	c = Context();
	c.launches = launches;
	return NBGPU_OK;
}

int nbgpu_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return count;
}

int nbgpu_device_info(char *buf, size_t len)
{
	NB_INIT();
	NB_ARG(buf != nullptr && len > 0);
	cudaDeviceProp prop;
	NB_CUDA(cudaGetDeviceProperties(&prop, g_ctx.device));
	int persist = 0, window = 0, smem_optin = 0;
	cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, g_ctx.device);
	cudaDeviceGetAttribute(&window, cudaDevAttrMaxAccessPolicyWindowSize, g_ctx.device);
	cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, g_ctx.device);
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	snprintf(buf, len,
		 "{\"name\": \"%s\", \"cc\": \"%d.%d\", \"sms\": %d, \"l2_bytes\": %d, "
		 "\"max_persisting_l2_bytes\": %d, \"max_access_policy_window_bytes\": %d, "
		 "\"smem_per_block_optin\": %d, \"hbm_total_bytes\": %zu, \"hbm_free_bytes\": %zu}",
		 prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.l2CacheSize, persist, window,
		 smem_optin, total_b, free_b);
	return NBGPU_OK;
}

int nbgpu_sync(void)
{
	NB_INIT();
	NB_CUDA(cudaStreamSynchronize(g_ctx.copy_stream));
	NB_CUDA(cudaStreamSynchronize(g_ctx.stream));
	return NBGPU_OK;
}

const char *nbgpu_last_error(void) { return g_error.c_str(); }

void *nbgpu_stream(void)
{
	if (ensure_init() != NBGPU_OK)
		return nullptr;
	return (void *)g_ctx.stream;
}

uint64_t nbgpu_launch_count(void)
{
	uint64_t n = g_default_ctx.launches;
	for (int d = 0; d < kMaxDeviceContexts; d++)
		n += g_device_ctx[d].launches;
	return n;
}

int nbgpu_thread_bind_device(int device)
{
	if (device < 0) {
		t_ctx = &g_default_ctx;
		return NBGPU_OK;
	}
	NB_ARG(device < kMaxDeviceContexts);
	t_ctx = &g_device_ctx[device];
	if (t_ctx->ready) {
		cudaSetDevice(device);
		return NBGPU_OK;
	}
	const int st = init_device(device);
	if (st != NBGPU_OK)
		t_ctx = &g_default_ctx;
	return st;
}

int nbgpu_malloc(void **d_ptr, size_t bytes)
{
	NB_INIT();
	NB_ARG(d_ptr != nullptr);
	cudaError_t e = nbgpu::dmalloc(d_ptr, bytes ? bytes : 1);
	if (e != cudaSuccess) {
		set_error("nbgpu::dmalloc(%zu): %s", bytes, cudaGetErrorString(e));
		cudaGetLastError();
		return NBGPU_ERR_NOMEM;
	}
	return NBGPU_OK;
}

int nbgpu_free(void *d_ptr)
{
	NB_INIT();
	if (d_ptr) {
		NB_CUDA(cudaStreamSynchronize(g_ctx.stream));
		NB_CUDA(nbgpu::dfree(d_ptr));
	}
	return NBGPU_OK;
}

int nbgpu_memset(void *d_ptr, int byte, size_t bytes)
{
	NB_INIT();
	NB_CUDA(cudaMemsetAsync(d_ptr, byte, bytes, g_ctx.stream));
	return NBGPU_OK;
}

int nbgpu_copy_h2d(void *d_dst, const void *src, size_t bytes)
{
	NB_INIT();
	NB_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
	NB_CUDA(cudaStreamSynchronize(g_ctx.stream));
	return NBGPU_OK;
}

int nbgpu_copy_d2h(void *dst, const void *d_src, size_t bytes)
{
	NB_INIT();
	NB_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
	NB_CUDA(cudaStreamSynchronize(g_ctx.stream));
	return NBGPU_OK;
}

int nbgpu_copy_d2d(void *d_dst, const void *d_src, size_t bytes)
{
	NB_INIT();
	NB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
	return NBGPU_OK;
}

int nbgpu_host_alloc(void **ptr, size_t bytes)
{
	NB_INIT();
	NB_ARG(ptr != nullptr);
	NB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
	return NBGPU_OK;
}

int nbgpu_host_free(void *ptr)
{
	NB_INIT();
	if (ptr)
		NB_CUDA(cudaFreeHost(ptr));
	return NBGPU_OK;
}

int nbgpu_timer_start(void)
{
	NB_INIT();
	NB_CUDA(cudaEventRecord(g_ctx.ev_a, g_ctx.stream));
	return NBGPU_OK;
}

int nbgpu_timer_stop(float *elapsed_ms)
{
	NB_INIT();
	NB_CUDA(cudaEventRecord(g_ctx.ev_b, g_ctx.stream));
	NB_CUDA(cudaEventSynchronize(g_ctx.ev_b));
	if (elapsed_ms)
		NB_CUDA(cudaEventElapsedTime(elapsed_ms, g_ctx.ev_a, g_ctx.ev_b));
	return NBGPU_OK;
}

}  // extern "C"
