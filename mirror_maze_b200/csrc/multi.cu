// multi.cu — mm_multi: the dispatch of reference src/main.rs:867-886 split over a list of GPUs inside ONE process, behind
// the C-ABI (SURVEY §8 b/e).  One mm_ctx per device (scene replicated), the frame's virtual groups interleaved over the
// devices, finished pixels exchanged by the render kernel's own NVLink peer stores, by an NCCL all-gather of tiles, or not
// at all (host assembly through zero-copy stores).  No torch, no second process: a C or Rust caller gets N GPUs from one
// call.  NCCL is resolved with dlopen only when MM_EXCHANGE_NCCL is asked for, so the library has no link-time NCCL
// dependency (and shares the process's libnccl when a host framework already loaded one).
#include <dlfcn.h>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "ctx.h"

using namespace mmk;
using namespace mmapi;

namespace {

// The slice of nccl.h this file uses (NCCL 2.x ABI): opaque communicator, result code 0 = success, ncclFloat32 = 7.
typedef void *nccl_comm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !AllGather || !GetErrorString) {
            err = "libnccl lacks an expected symbol";
            return false;
        }
        return true;
    }
};
constexpr int kNcclFloat32 = 7;

thread_local std::string g_multi_create_err;

}  // namespace

struct mm_multi {
    int n = 0, exchange = MM_EXCHANGE_PEER;
    std::vector<int> dev;
    std::vector<mm_ctx *> ctx;
    std::vector<float *> frames;          // per device: H*W*4 floats
    std::vector<float *> tiles, gathered; // MM_EXCHANGE_NCCL: own tiles [max_count][ppc][4], all tiles [n*max_count][ppc][4]
    size_t tiles_cap = 0;                 // floats per device in `tiles`
    uint32_t fw = 0, fh = 0;
    std::vector<nccl_comm_t> comms;
    NcclApi nccl;
    std::string err;
    // staged host output (caller buffer not pinned): device 0's frame -> pinned staging -> caller
    float *h_stage = nullptr;
    size_t stage_bytes = 0;
    float *pending_out = nullptr;
    size_t pending_bytes = 0;
    bool in_flight = false;
    std::vector<char> launched;           // which devices had groups in the frame in flight
};

namespace {

int mfail(mm_multi *m, int code, const std::string &msg) {
    m->err = msg;
    return code;
}
int child_fail(mm_multi *m, int i, int rc) {
    m->err = "device " + std::to_string(m->dev[i]) + ": " + mm_last_error(m->ctx[i]);
    return rc;
}
#define MCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            m->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                 \
            return MM_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

void free_frames(mm_multi *m) {
    for (int i = 0; i < m->n; i++) {
        cudaSetDevice(m->dev[i]);
        if (i < (int)m->frames.size()) { cudaFree(m->frames[i]); m->frames[i] = nullptr; }
        if (i < (int)m->tiles.size()) { cudaFree(m->tiles[i]); m->tiles[i] = nullptr; }
        if (i < (int)m->gathered.size()) { cudaFree(m->gathered[i]); m->gathered[i] = nullptr; }
    }
    m->fw = m->fh = 0;
    m->tiles_cap = 0;
}

int ensure_frames(mm_multi *m, uint32_t W, uint32_t H, size_t tile_floats) {
    if (m->fw != W || m->fh != H) {
        for (int i = 0; i < m->n; i++) { cudaSetDevice(m->dev[i]); cudaStreamSynchronize(m->ctx[i]->stream); }
        free_frames(m);
        const size_t bytes = (size_t)W * H * 4 * sizeof(float);
        for (int i = 0; i < m->n; i++) {
            MCK(cudaSetDevice(m->dev[i]));
            MCK(cudaMalloc(&m->frames[i], bytes));
            MCK(cudaMemsetAsync(m->frames[i], 0, bytes, m->ctx[i]->stream));
        }
        for (int i = 0; i < m->n; i++) { MCK(cudaSetDevice(m->dev[i])); MCK(cudaStreamSynchronize(m->ctx[i]->stream)); }
        m->fw = W; m->fh = H;
    }
    if (m->exchange == MM_EXCHANGE_NCCL && m->tiles_cap < tile_floats) {
        for (int i = 0; i < m->n; i++) {
            MCK(cudaSetDevice(m->dev[i]));
            MCK(cudaStreamSynchronize(m->ctx[i]->stream));
            cudaFree(m->tiles[i]); cudaFree(m->gathered[i]);
            m->tiles[i] = m->gathered[i] = nullptr;
            MCK(cudaMalloc(&m->tiles[i], tile_floats * sizeof(float)));
            MCK(cudaMemsetAsync(m->tiles[i], 0, tile_floats * sizeof(float), m->ctx[i]->stream));
            MCK(cudaMalloc(&m->gathered[i], tile_floats * sizeof(float) * (size_t)m->n));
        }
        m->tiles_cap = tile_floats;
    }
    return MM_OK;
}

}  // namespace

extern "C" {

int mm_multi_create(const int *cuda_devices, int n_devices, int exchange, mm_multi **out) {
    if (!out) return MM_ERR_INVALID;
    *out = nullptr;
    if (!cuda_devices || n_devices < 1 || n_devices > MM_MAX_PEERS) { g_multi_create_err = "1..MM_MAX_PEERS devices"; return MM_ERR_INVALID; }
    if (exchange != MM_EXCHANGE_PEER && exchange != MM_EXCHANGE_NCCL && exchange != MM_EXCHANGE_NONE) {
        g_multi_create_err = "unknown exchange mode";
        return MM_ERR_INVALID;
    }
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (cuda_devices[i] == cuda_devices[j]) { g_multi_create_err = "device listed twice"; return MM_ERR_INVALID; }
    mm_multi *m = new (std::nothrow) mm_multi();
    if (!m) return MM_ERR_NOMEM;
    m->n = n_devices;
    m->exchange = n_devices == 1 ? MM_EXCHANGE_NONE : exchange;
    m->dev.assign(cuda_devices, cuda_devices + n_devices);
    m->ctx.assign(n_devices, nullptr);
    m->frames.assign(n_devices, nullptr);
    m->tiles.assign(n_devices, nullptr);
    m->gathered.assign(n_devices, nullptr);
    m->launched.assign(n_devices, 0);
    for (int i = 0; i < n_devices; i++) {
        int rc = mm_create(cuda_devices[i], &m->ctx[i]);
        if (rc != MM_OK) {
            g_multi_create_err = std::string("device ") + std::to_string(cuda_devices[i]) + ": " + mm_last_error(nullptr);
            mm_multi_destroy(m);
            return rc;
        }
    }
    if (m->exchange == MM_EXCHANGE_PEER) {
        for (int i = 0; i < n_devices; i++)
            for (int j = 0; j < n_devices; j++) {
                if (i == j) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, cuda_devices[i], cuda_devices[j]);
                cudaError_t e = cudaErrorPeerAccessUnsupported;
                if (can) {
                    cudaSetDevice(cuda_devices[i]);
                    e = cudaDeviceEnablePeerAccess(cuda_devices[j], 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                }
                if (e != cudaSuccess) {
                    g_multi_create_err = "no peer access from device " + std::to_string(cuda_devices[i]) + " to " + std::to_string(cuda_devices[j]) +
                                         " (" + cudaGetErrorString(e) + "): use MM_EXCHANGE_NCCL or MM_EXCHANGE_NONE";
                    cudaGetLastError();
                    mm_multi_destroy(m);
                    return MM_ERR_UNSUPPORTED;
                }
            }
    } else if (m->exchange == MM_EXCHANGE_NCCL) {
        std::string e;
        if (!m->nccl.load(e)) { g_multi_create_err = e; mm_multi_destroy(m); return MM_ERR_UNSUPPORTED; }
        m->comms.assign(n_devices, nullptr);
        int r = m->nccl.CommInitAll(m->comms.data(), n_devices, cuda_devices);
        if (r != 0) {
            g_multi_create_err = std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(r);
            m->comms.clear();
            mm_multi_destroy(m);
            return MM_ERR_CUDA;
        }
    }
    g_multi_create_err.clear();
    *out = m;
    return MM_OK;
}

int mm_multi_destroy(mm_multi *m) {
    if (!m) return MM_OK;
    for (int i = 0; i < m->n; i++)
        if (m->ctx[i]) { cudaSetDevice(m->dev[i]); cudaStreamSynchronize(m->ctx[i]->stream); }
    for (nccl_comm_t c : m->comms)
        if (c) m->nccl.CommDestroy(c);
    free_frames(m);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    for (mm_ctx *c : m->ctx) mm_destroy(c);
    delete m;
    return MM_OK;
}

const char *mm_multi_last_error(const mm_multi *m) { return m ? m->err.c_str() : g_multi_create_err.c_str(); }
int mm_multi_n_devices(const mm_multi *m) { return m ? m->n : 0; }
mm_ctx *mm_multi_ctx(mm_multi *m, int index) { return (m && index >= 0 && index < m->n) ? m->ctx[index] : nullptr; }

int mm_multi_upload_scene(mm_multi *m, const mm_plane *planes, uint32_t n_planes, const mm_bvh_node *nodes, uint32_t n_nodes,
                          const uint32_t *indices, const uint8_t *materials, const mm_float4 *emissions, const uint8_t *noise_rgba8,
                          uint32_t noise_w, uint32_t noise_h) {
    if (!m) return MM_ERR_INVALID;
    m->err.clear();
    for (int i = 0; i < m->n; i++) {   // replicated: <= 9 MB per device (SURVEY §8 e)
        int rc = mm_upload_scene(m->ctx[i], planes, n_planes, nodes, n_nodes, indices, materials, emissions, noise_rgba8, noise_w, noise_h);
        if (rc != MM_OK) return child_fail(m, i, rc);
    }
    return MM_OK;
}

int mm_multi_wait(mm_multi *m, mm_counters *counters) {
    if (!m) return MM_ERR_INVALID;
    m->err.clear();
    for (int i = 0; i < m->n; i++) {
        MCK(cudaSetDevice(m->dev[i]));
        MCK(cudaStreamSynchronize(m->ctx[i]->stream));
    }
    if (m->pending_out) {
        memcpy(m->pending_out, m->h_stage, m->pending_bytes);
        m->pending_out = nullptr;
    }
    m->in_flight = false;
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (int i = 0; i < m->n; i++) {
            if (!m->launched[i]) continue;
            mm_counters c;
            counters_out(m->ctx[i]->h_counters, &c);
            counters->paths += c.paths; counters->rays += c.rays; counters->inner_visits += c.inner_visits;
            counters->leaf_visits += c.leaf_visits; counters->rect_tests += c.rect_tests; counters->hits += c.hits;
            counters->literal_rays += c.literal_rays;
            if (c.max_stack > counters->max_stack) counters->max_stack = c.max_stack;
        }
    }
    return MM_OK;
}

int mm_multi_render_async(mm_multi *m, const mm_uniform *uni, const mm_params *params, const mm_chunk *chunks, uint32_t n_chunks,
                          float *out_rgba) {
    if (!m) return MM_ERR_INVALID;
    m->err.clear();
    if (!uni || !params) return mfail(m, MM_ERR_INVALID, "null uniform or params");
    int rc;
    if (m->in_flight && (rc = mm_multi_wait(m, nullptr)) != MM_OK) return rc;
    const int n = m->n;
    if (chunks) {
        for (int i = 0; i < n; i++)
            if ((rc = mm_set_chunks(m->ctx[i], chunks, n_chunks)) != MM_OK) return child_fail(m, i, rc);
    } else if (m->ctx[0]->n_chunks == 0) {
        return mfail(m, MM_ERR_INVALID, "null chunk list and none set earlier");
    }
    const uint64_t n_groups = (uint64_t)params->grid_x * params->grid_y;
    if (n_groups == 0 || n_groups > 0xFFFFFFFFull) return mfail(m, MM_ERR_INVALID, "empty or oversized grid");
    // interleaved partition (tile_partition in renderer.py): device i renders groups i, i + n, ...
    std::vector<Launch> L(n);
    uint32_t max_count = 0;
    for (int i = 0; i < n; i++) {
        mm_params q = *params;
        q.group_first = (uint32_t)i; q.group_step = (uint32_t)n;
        q.group_count = n_groups > (uint64_t)i ? (uint32_t)((n_groups - i + n - 1) / n) : 0;
        if (q.group_count > max_count) max_count = q.group_count;
        m->launched[i] = q.group_count ? 1 : 0;
        if (!q.group_count) continue;
        MCK(cudaSetDevice(m->dev[i]));
        if ((rc = build_launch(m->ctx[i], uni, &q, false, L[i])) != MM_OK) return child_fail(m, i, rc);
    }
    const KParams &p0 = L[0].p;    // device 0 always owns group 0
    const uint32_t W = p0.W, H = p0.H, ppc = p0.ppc;
    const size_t frame_bytes = (size_t)W * H * 4 * sizeof(float);
    if ((rc = ensure_frames(m, W, H, (size_t)max_count * ppc * 4)) != MM_OK) return rc;
    void *pinned = out_rgba ? host_device_alias(out_rgba, frame_bytes) : nullptr;
    void *alias = (params->flags & MM_FLAG_NO_ZERO_COPY) ? nullptr : pinned;
    if (out_rgba && !alias && m->exchange == MM_EXCHANGE_NONE && n > 1)
        return mfail(m, MM_ERR_INVALID, "MM_EXCHANGE_NONE assembles the frame in host memory: out_rgba must be mapped pinned memory (mm_host_alloc / mm_host_register)");
    const bool staged = out_rgba && !pinned;
    if (staged && m->stage_bytes < frame_bytes) {
        if (m->h_stage) cudaFreeHost(m->h_stage);
        m->h_stage = nullptr; m->stage_bytes = 0;
        MCK(cudaMallocHost(&m->h_stage, frame_bytes));
        m->stage_bytes = frame_bytes;
    }
    // launch every device's share
    for (int i = 0; i < n; i++) {
        if (!m->launched[i]) continue;
        MCK(cudaSetDevice(m->dev[i]));
        KParams &p = L[i].p;
        p.host_out = static_cast<float *>(alias);
        if (m->exchange == MM_EXCHANGE_PEER) {
            p.n_peers = (uint32_t)n;
            for (int j = 0; j < n; j++) p.peers[j] = m->frames[j];
        } else if (m->exchange == MM_EXCHANGE_NCCL) {
            p.tiles = m->tiles[i];
        } else {
            p.image = m->frames[i];
        }
        if ((rc = do_launch(m->ctx[i], L[i])) != MM_OK) return child_fail(m, i, rc);
    }
    m->in_flight = true;
    if (m->exchange == MM_EXCHANGE_NCCL) {
        const size_t count = (size_t)max_count * ppc * 4;
        int r = m->nccl.GroupStart();
        for (int i = 0; i < n && r == 0; i++)
            r = m->nccl.AllGather(m->tiles[i], m->gathered[i], count, kNcclFloat32, m->comms[i], m->ctx[i]->stream);
        const int r2 = m->nccl.GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return mfail(m, MM_ERR_CUDA, std::string("ncclAllGather: ") + m->nccl.GetErrorString(r));
        for (int i = 0; i < n; i++) {
            MCK(cudaSetDevice(m->dev[i]));
            MCK(launch_scatter_all(m->gathered[i], m->frames[i], m->ctx[i]->d_chunks, (uint32_t)n, max_count, (uint32_t)n_groups,
                                   uni->chunk_width, W, H, m->ctx[i]->stream));
        }
    }
    if (out_rgba && !alias) {
        // device 0's assembled frame.  PEER: its kernel's end does not imply the other devices' stores have landed, so its
        // stream first waits for every device's kernel-end event.
        MCK(cudaSetDevice(m->dev[0]));
        if (m->exchange == MM_EXCHANGE_PEER)
            for (int j = 1; j < n; j++)
                if (m->launched[j]) MCK(cudaStreamWaitEvent(m->ctx[0]->stream, m->ctx[j]->ev1, 0));
        float *dst = staged ? m->h_stage : out_rgba;
        MCK(cudaMemcpyAsync(dst, m->frames[0], frame_bytes, cudaMemcpyDeviceToHost, m->ctx[0]->stream));
        if (staged) { m->pending_out = out_rgba; m->pending_bytes = frame_bytes; }
    }
    return MM_OK;
}

int mm_multi_render(mm_multi *m, const mm_uniform *uni, const mm_params *params, const mm_chunk *chunks, uint32_t n_chunks,
                    float *out_rgba, mm_counters *counters) {
    int rc = mm_multi_render_async(m, uni, params, chunks, n_chunks, out_rgba);
    if (rc != MM_OK) return rc;
    return mm_multi_wait(m, counters);
}

int mm_multi_frame_device(mm_multi *m, int index, float **d_frame) {
    if (!m || !d_frame || index < 0 || index >= m->n) return MM_ERR_INVALID;
    if (!m->frames[index]) return mfail(m, MM_ERR_INVALID, "mm_multi_frame_device: nothing rendered yet");
    *d_frame = m->frames[index];
    return MM_OK;
}

int mm_multi_last_ms(mm_multi *m, float *ms) {
    if (!m || !ms) return MM_ERR_INVALID;
    m->err.clear();
    float worst = 0.0f;
    for (int i = 0; i < m->n; i++) {
        if (!m->launched[i]) continue;
        float t = 0.0f;
        int rc = mm_last_ms(m->ctx[i], &t);
        if (rc != MM_OK) return child_fail(m, i, rc);
        if (t > worst) worst = t;
    }
    *ms = worst;
    return MM_OK;
}

}  // extern "C"
