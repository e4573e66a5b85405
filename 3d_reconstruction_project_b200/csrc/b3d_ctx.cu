// b3d_ctx.cu -- context, error text, stream-ordered scratch memory.
#include "b3d_common.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace b3d {

static thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

}  // namespace b3d

namespace {
// B3D_TRACE_SLOW=<ms>: report host-side runtime calls that take longer than <ms> (allocator growth, stalled syncs)
double trace_slow_ms() {
    static const double v = [] {
        const char* e = getenv("B3D_TRACE_SLOW");
        return e ? atof(e) : 0.0;
    }();
    return v;
}
struct SlowCall {
    const char* what;
    size_t bytes;
    std::chrono::steady_clock::time_point t0;
    SlowCall(const char* w, size_t b) : what(w), bytes(b), t0(std::chrono::steady_clock::now()) {}
    ~SlowCall() {
        const double lim = trace_slow_ms();
        if (lim <= 0) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > lim) fprintf(stderr, "[b3d slow] %s(%zu) took %.1f ms\n", what, bytes, ms);
    }
};
}  // namespace

int b3d_ctx::bind() const {
    B3D_CUDA(cudaSetDevice(device));
    return B3D_OK;
}
int b3d_ctx::alloc_bytes(void** p, size_t bytes) {
    SlowCall sc("alloc", bytes);
    const size_t gran = bytes < (1u << 20) ? 512 : (2u << 20);
    const size_t want = (bytes + gran - 1) / gran * gran;
    auto it = cache_free.lower_bound(want);
    if (it != cache_free.end() && it->first <= want + want / 2 + (1u << 20)) {
        *p = it->second;
        cache_live[*p] = it->first;
        cache_free_bytes -= it->first;
        cache_free.erase(it);
        return B3D_OK;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaErrorMemoryAllocation) {
        // give the cached blocks back to the driver and retry once
        cudaGetLastError();
        cudaStreamSynchronize(stream);
        for (auto& kv : cache_free) {
            cudaFree(kv.second);
            cache_total -= kv.first;
        }
        cache_free.clear();
        cache_free_bytes = 0;
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        cudaGetLastError();
        return b3d::set_error(e == cudaErrorMemoryAllocation ? B3D_E_NOMEM : B3D_E_CUDA, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e));
    }
    cache_live[*p] = want;
    cache_total += want;
    return B3D_OK;
}
void b3d_ctx::free_async(void* p) {
    if (!p) return;
    auto it = cache_live.find(p);
    if (it == cache_live.end()) return;
    cache_free.emplace(it->second, p);
    cache_free_bytes += it->second;
    cache_live.erase(it);
    // Workloads whose sizes keep changing (a growing map, different batch sizes) would let the free list grow without bound:
    // above the cap the idle blocks go back to the driver (after the stream drained; the next step re-allocates what it needs).
    if (cache_free_bytes > cache_cap_bytes) {
        cudaStreamSynchronize(stream);
        for (auto& kv : cache_free) {
            cudaFree(kv.second);
            cache_total -= kv.first;
        }
        cache_free.clear();
        cache_free_bytes = 0;
    }
}
void b3d_ctx::cache_release_all() {
    cudaStreamSynchronize(stream);
    for (auto& kv : cache_free) cudaFree(kv.second);
    for (auto& kv : cache_live) cudaFree(kv.first);
    cache_free.clear();
    cache_live.clear();
    cache_total = 0;
    cache_free_bytes = 0;
}
int b3d_ctx::sync() {
    SlowCall sc("cudaStreamSynchronize", 0);
    B3D_CUDA(cudaStreamSynchronize(stream));
    return B3D_OK;
}
int b3d_ctx::ensure_pinned(size_t bytes) {
    if (bytes <= pinned_bytes) return B3D_OK;
    // the old block may still be the target of an in-flight copy
    B3D_CUDA(cudaStreamSynchronize(stream));
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_bytes = 0;
    size_t want = 4096;
    while (want < bytes) want <<= 1;
    cudaError_t e = cudaMallocHost(&pinned, want);
    if (e != cudaSuccess) return b3d::set_error(B3D_E_NOMEM, "cudaMallocHost(%zu): %s", want, cudaGetErrorString(e));
    pinned_bytes = want;
    return B3D_OK;
}
int b3d_ctx::download(void* dst_h, const void* src_d, size_t bytes) {
    if (bytes == 0) return B3D_OK;
    B3D_TRY(ensure_pinned(bytes));
    SlowCall sc("download", bytes);
    B3D_CUDA(cudaMemcpyAsync(pinned, src_d, bytes, cudaMemcpyDeviceToHost, stream));
    B3D_CUDA(cudaStreamSynchronize(stream));
    memcpy(dst_h, pinned, bytes);
    return B3D_OK;
}
int b3d_ctx::upload(void* dst_d, const void* src_h, size_t bytes) {
    if (bytes == 0) return B3D_OK;
    SlowCall sc("upload", bytes);
    B3D_CUDA(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, stream));
    return B3D_OK;
}

void b3d_ctx::prof_begin(const char* name) {
    ProfRec r{name, nullptr, nullptr, 0};
    for (cudaEvent_t* e : {&r.e0, &r.e1}) {
        if (!prof_pool.empty()) {
            *e = prof_pool.back();
            prof_pool.pop_back();
        } else {
            cudaEventCreate(e);
        }
    }
    cudaEventRecord(r.e0, stream);
    prof.push_back(r);
}
void b3d_ctx::prof_end() {
    if (!prof.empty()) cudaEventRecord(prof.back().e1, stream);
}

extern "C" {

int b3d_version(void) { return 100; }

int b3d_ctx_profile(b3d_ctx* ctx, int enable) {
    if (!ctx) return b3d::set_error(B3D_E_INVALID, "ctx is NULL");
    B3D_TRY(ctx->bind());
    B3D_TRY(ctx->sync());
    for (auto& r : ctx->prof) {
        ctx->prof_pool.push_back(r.e0);
        ctx->prof_pool.push_back(r.e1);
    }
    ctx->prof.clear();
    ctx->profiling = enable != 0;
    return B3D_OK;
}

// Writes one line per kernel name: "<name>\t<launches>\t<total_ms>\n" (sorted by total time, descending) and clears the
// records. Returns the number of bytes needed (excluding the terminator); call with cap = 0 to size the buffer.
int64_t b3d_ctx_profile_report(b3d_ctx* ctx, char* buf, int64_t cap) {
    if (!ctx) return -1;
    if (ctx->bind() != B3D_OK || ctx->sync() != B3D_OK) return -1;
    struct Agg {
        std::string name;
        int64_t n;
        double ms;
        int64_t bytes;
    };
    std::vector<Agg> agg;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = 0.f;
        std::string nm(r.name);
        size_t lt = nm.find('<');  // strip template arguments and parentheses: "(normals_kernel<T, 32, TENSOR>)" -> "normals_kernel"
        if (lt != std::string::npos) nm = nm.substr(0, lt);
        while (!nm.empty() && (nm.front() == '(' || nm.front() == ' ')) nm.erase(nm.begin());
        while (!nm.empty() && (nm.back() == ')' || nm.back() == ' ')) nm.pop_back();
        bool found = false;
        for (auto& a : agg)
            if (a.name == nm) {
                a.n += 1;
                a.ms += ms;
                a.bytes += r.bytes;
                found = true;
                break;
            }
        if (!found) agg.push_back({nm, 1, (double)ms, r.bytes});
    }
    for (size_t i = 0; i < agg.size(); ++i)
        for (size_t j = i + 1; j < agg.size(); ++j)
            if (agg[j].ms > agg[i].ms) std::swap(agg[i], agg[j]);
    std::string out;
    char line[256];
    for (auto& a : agg) {
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%lld\n", a.name.c_str(), (long long)a.n, a.ms, (long long)a.bytes);
        out += line;
    }
    if (buf && cap > 0) {
        size_t m = std::min<size_t>(out.size(), (size_t)cap - 1);
        memcpy(buf, out.data(), m);
        buf[m] = 0;
        if ((int64_t)out.size() < cap) {
            for (auto& r : ctx->prof) {
                ctx->prof_pool.push_back(r.e0);
                ctx->prof_pool.push_back(r.e1);
            }
            ctx->prof.clear();
        }
    }
    return (int64_t)out.size();
}

const char* b3d_last_error(void) { return b3d::g_last_error.c_str(); }

int b3d_ctx_create(int device, void* stream, b3d_ctx** out) {
    if (!out) return b3d::set_error(B3D_E_INVALID, "b3d_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return b3d::set_error(B3D_E_CUDA, "b3d_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return b3d::set_error(B3D_E_INVALID, "b3d_ctx_create: device %d out of range (0..%d)", device, count - 1);
    B3D_CUDA(cudaSetDevice(device));
    b3d_ctx* c = new b3d_ctx();
    c->device = device;
    // NULL is the legacy default stream (what torch.cuda.current_stream() is unless the caller switched streams), so the
    // kernels stay ordered with the caller's own copies and allocations on that stream
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    c->own_stream = false;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
        c->cache_cap_bytes = prop.totalGlobalMem / 4;  // idle scratch kept for re-use: at most a quarter of the device memory
    }
    c->pinned_bytes = 4096;
    if (cudaMallocHost(&c->pinned, c->pinned_bytes) != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return b3d::set_error(B3D_E_NOMEM, "cudaMallocHost failed");
    }
    *out = c;
    return B3D_OK;
}

int b3d_ctx_destroy(b3d_ctx* ctx) {
    if (!ctx) return B3D_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->cache_release_all();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (auto& r : ctx->prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B3D_OK;
}

int b3d_ctx_synchronize(b3d_ctx* ctx) {
    if (!ctx) return b3d::set_error(B3D_E_INVALID, "ctx is NULL");
    B3D_TRY(ctx->bind());
    return ctx->sync();
}

int64_t b3d_ctx_launch_count(b3d_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
