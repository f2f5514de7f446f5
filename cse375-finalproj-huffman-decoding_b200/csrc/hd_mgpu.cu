// hd_mgpu.cu -- multi-GPU sharded decode behind the C ABI (hd_mgpu_*, include/hd_b200.h), both schemes.
//
// Replaces the partition / launch / stitch code of gpuhd/multigpu_demo.cc:178-313 (equal split of units AND symbols:
// incorrect by construction, README.md:37 "TESTS FAIL").  One hd_mgpu object per rank (one process per GPU, or one
// host thread per GPU of a single process).  The stream is cut on 128-bit segment boundaries (whole 256-segment
// tiles); every rank decodes its shard from local HBM into a local buffer; the only exchange is three integers per
// rank {symbols decoded, entry offset (| unconverged << 32), exit offset}, all-gathered through a peer-memory
// mailbox (hd_mailbox_*), a caller-supplied ncclComm_t, or a caller-supplied callback.  Built only on the public
// single-GPU entry points of the library.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "hd_b200.h"

namespace {

constexpr uint64_t kSegUnits = 4, kDefaultHaloSegs = 4096;
constexpr uint64_t kEmpty = ~0ull;          // entry / exit of a rank without segments
constexpr uint64_t kUnconvergedBit = 1ull << 32;
constexpr int kFlagRing = 256;

inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// ncclAllGather of the NCCL that is already loaded in the process (the caller owns a communicator, so it is); the
// library itself does not link against NCCL
typedef int (*nccl_allgather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
nccl_allgather_t find_nccl_allgather() {
    static const nccl_allgather_t fn = [] {   // (resolved once; thread-safe: one host thread per GPU may get here at the same time)
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW);
            if (h) sym = dlsym(h, "ncclAllGather");
        }
        return reinterpret_cast<nccl_allgather_t>(sym);
    }();
    return fn;
}
constexpr int kNcclUint64 = 5;   // ncclDataType_t ncclUint64 (nccl.h)

}  // namespace

struct hd_mgpu {
    hd_scheme scheme = HD_SCHEME_SELFSYNC;
    uint32_t rank = 0, world = 1;
    hd_mgpu_plan plan = {};
    // bound buffers
    const uint32_t *d_units = nullptr;
    const uint32_t *d_gap_words = nullptr;
    const hd_table *table = nullptr;
    uint8_t *d_out = nullptr;
    void *d_ws = nullptr;
    // exchange
    void *nccl_comm = nullptr;
    hd_mailbox *mailbox = nullptr;
    hd_allgather_fn callback = nullptr;
    void *callback_user = nullptr;
    uint64_t seq = 0;
    // device / pinned state
    uint64_t *d_info = nullptr;     // {count, entry | unconverged << 32, exit}
    uint64_t *d_all = nullptr;      // world x 3
    uint32_t *h_flags = nullptr;    // pinned ring of convergence flags (one per enqueued step)
    uint64_t *h_info = nullptr;     // pinned: the triple of a host-buffer step (hd_mgpu_host_step)
    int next_flag = 0;
    bool pending = false;
    uint32_t repairs = 0;           // decodes repeated by the last hd_mgpu_finish (unconverged + seam repairs)
    std::vector<uint64_t> infos;    // host copy of the last exchange
};

#define MG_CUDA(call)                      \
    do {                                   \
        if ((call) != cudaSuccess) {       \
            (void)cudaGetLastError();      \
            return HD_ERR_CUDA;            \
        }                                  \
    } while (0)

extern "C" {

// ------------------------------------------------------------------------------------------------ pure host part
hd_status hd_mgpu_plan_shard(hd_scheme scheme, uint64_t n_units, uint32_t rank, uint32_t world,
                             uint32_t min_code_length, uint64_t halo_segments, hd_mgpu_plan *plan) {
    if (!plan || world == 0 || rank >= world || n_units == 0) return HD_ERR_INVALID;
    if (scheme != HD_SCHEME_SELFSYNC && scheme != HD_SCHEME_GAP) return HD_ERR_INVALID;
    if (min_code_length == 0) min_code_length = 1;
    if (halo_segments == 0) halo_segments = kDefaultHaloSegs;
    std::memset(plan, 0, sizeof(*plan));
    const uint64_t n_seg = div_up(n_units, kSegUnits);
    hd_shard_segments(n_seg, rank, world, &plan->seg_begin, &plan->seg_end);
    const uint64_t b = plan->seg_begin, e = plan->seg_end;
    if (b >= e) {   // fewer tiles than ranks: this rank has nothing to decode
        plan->seg_end = b;
        plan->first_unit = std::min(b * kSegUnits, n_units);
        return HD_OK;
    }
    uint64_t halo = 0;
    if (scheme == HD_SCHEME_SELFSYNC && rank > 0) halo = std::min(halo_segments, b) / 8 * 8;   // whole gap words
    const uint64_t end_unit = std::min(e * kSegUnits, n_units);
    plan->halo_units = halo * kSegUnits;
    plan->first_unit = (b - halo) * kSegUnits;
    plan->shard_units = end_unit - b * kSegUnits;
    plan->readable_units = std::min(end_unit + kSegUnits, n_units) - plan->first_unit;   // right halo: the last codeword may end there
    if (scheme == HD_SCHEME_GAP) {
        // elements seg_begin-1 (the shard's entry offset: top nibble of the word before) .. seg_end-1
        plan->first_gap_word = b ? b / 8 - 1 : 0;
        plan->n_gap_words = div_up(e, 8) - plan->first_gap_word;
        plan->workspace_bytes = hd_workspace_bytes_gap(plan->readable_units);
    } else {
        plan->workspace_bytes = hd_workspace_bytes_selfsync(plan->halo_units + plan->shard_units);
    }
    plan->out_capacity = (plan->shard_units * 32 + 15 + min_code_length - 1) / min_code_length + 64;
    return HD_OK;
}

// first rank whose assumed entry offset differs from the offset at which its nearest non-empty predecessor left its
// shard (world = every seam is consistent)
uint32_t hd_mgpu_first_bad_seam(const uint64_t *infos, uint32_t world) {
    uint64_t prev_exit = kEmpty;
    for (uint32_t g = 0; g < world; ++g) {
        const uint64_t entry = infos[3 * g + 1], exit = infos[3 * g + 2];
        if (entry == kEmpty) continue;
        if (prev_exit != kEmpty && (entry & 0xffffu) != (prev_exit & 0xffffu)) return g;
        prev_exit = exit;
    }
    return world;
}

// exclusive scan of the counts, clamped to the stream's symbol count (the last shard also decodes the pad bits)
void hd_mgpu_place(const uint64_t *infos, uint32_t world, uint64_t n_symbols, uint32_t rank, uint64_t *offset,
                   uint64_t *count) {
    uint64_t off = 0;
    for (uint32_t g = 0; g < rank && g < world; ++g) off += infos[3 * g];
    const uint64_t c = rank < world ? infos[3 * rank] : 0;
    if (offset) *offset = std::min(off, n_symbols);
    if (count) *count = off >= n_symbols ? 0 : std::min(c, n_symbols - off);
}

// ------------------------------------------------------------------------------------------------------ context
hd_status hd_mgpu_create(hd_scheme scheme, uint32_t rank, uint32_t world, const hd_mgpu_plan *plan, hd_mgpu **out) {
    if (!plan || !out || world == 0 || rank >= world) return HD_ERR_INVALID;
    hd_mgpu *m = new (std::nothrow) hd_mgpu();
    if (!m) return HD_ERR_INVALID;
    m->scheme = scheme;
    m->rank = rank;
    m->world = world;
    m->plan = *plan;
    m->infos.assign(3 * (size_t)world, 0);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&m->d_info), 3 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_all), 3 * sizeof(uint64_t) * world);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->h_flags), kFlagRing * sizeof(uint32_t), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->h_info), 4 * sizeof(uint64_t), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMemset(m->d_info, 0, 3 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemset(m->d_all, 0, 3 * sizeof(uint64_t) * world);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        hd_mgpu_destroy(m);
        return HD_ERR_CUDA;
    }
    std::memset(m->h_flags, 0, kFlagRing * sizeof(uint32_t));
    *out = m;
    return HD_OK;
}

void hd_mgpu_destroy(hd_mgpu *m) {
    if (!m) return;
    if (m->d_info) cudaFree(m->d_info);
    if (m->d_all) cudaFree(m->d_all);
    if (m->h_flags) cudaFreeHost(m->h_flags);
    if (m->h_info) cudaFreeHost(m->h_info);
    delete m;
}

hd_status hd_mgpu_exchange_nccl(hd_mgpu *m, void *nccl_comm) {
    if (!m || !nccl_comm) return HD_ERR_INVALID;
    if (!find_nccl_allgather()) return HD_ERR_INVALID;
    m->nccl_comm = nccl_comm;
    m->mailbox = nullptr;
    m->callback = nullptr;
    return HD_OK;
}
hd_status hd_mgpu_exchange_mailbox(hd_mgpu *m, hd_mailbox *mb) {
    if (!m || !mb) return HD_ERR_INVALID;
    m->mailbox = mb;
    m->nccl_comm = nullptr;
    m->callback = nullptr;
    return HD_OK;
}
hd_status hd_mgpu_exchange_callback(hd_mgpu *m, hd_allgather_fn fn, void *user) {
    if (!m || !fn) return HD_ERR_INVALID;
    m->callback = fn;
    m->callback_user = user;
    m->mailbox = nullptr;
    m->nccl_comm = nullptr;
    return HD_OK;
}

hd_status hd_mgpu_bind(hd_mgpu *m, const uint32_t *d_units, const uint32_t *d_gap_words, const hd_table *table,
                       uint8_t *d_out, void *d_workspace) {
    if (!m || !table) return HD_ERR_INVALID;
    if (m->plan.shard_units && (!d_units || !d_out || !d_workspace)) return HD_ERR_INVALID;
    if (m->plan.shard_units && m->scheme == HD_SCHEME_GAP && !d_gap_words) return HD_ERR_INVALID;
    m->d_units = d_units;
    m->d_gap_words = d_gap_words;
    m->table = table;
    m->d_out = d_out;
    m->d_ws = d_workspace;
    return HD_OK;
}

// all-gather of d_info into d_all on `stream` (collective: every rank calls it the same number of times)
static hd_status mg_exchange(hd_mgpu *m, cudaStream_t s) {
    if (m->world == 1) {
        MG_CUDA(cudaMemcpyAsync(m->d_all, m->d_info, 3 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        return HD_OK;
    }
    if (m->mailbox) {
        m->seq += 1;
        hd_status st = hd_mailbox_publish(m->mailbox, m->d_info, m->seq, s);
        if (st != HD_OK) return st;
        return hd_mailbox_wait(m->mailbox, m->seq, m->d_all, s);
    }
    if (m->nccl_comm) {
        nccl_allgather_t ag = find_nccl_allgather();
        if (!ag) return HD_ERR_INVALID;
        return ag(m->d_info, m->d_all, 3, kNcclUint64, m->nccl_comm, s) == 0 ? HD_OK : HD_ERR_CUDA;
    }
    if (m->callback) return m->callback(m->d_info, m->d_all, 3 * sizeof(uint64_t), s, m->callback_user) == 0 ? HD_OK : HD_ERR_CUDA;
    return HD_ERR_INVALID;   // world > 1 needs an exchange
}

// the shard decode proper; first_bit >= 0: repair from an exact entry offset (self-sync: without the halo)
static hd_status mg_decode(hd_mgpu *m, int exact_first_bit, bool synchronous, cudaStream_t s) {
    const hd_mgpu_plan &p = m->plan;
    if (p.shard_units == 0) {   // empty shard: publish {0, empty, empty}
        const uint64_t h[3] = {0, kEmpty, kEmpty};
        MG_CUDA(cudaMemcpyAsync(m->d_info, h, sizeof(h), cudaMemcpyHostToDevice, s));
        MG_CUDA(cudaStreamSynchronize(s));   // (h is on this stack frame; an empty shard has nothing to overlap)
        return HD_OK;
    }
    if (m->scheme == HD_SCHEME_GAP) {
        const bool from_prev = p.seg_begin != 0;
        return hd_decode_gap_shard(m->d_units, p.readable_units, m->d_gap_words + (from_prev ? 1 : 0), p.seg_end - p.seg_begin,
                                   from_prev ? 1 : 0, m->table, m->d_out, p.out_capacity, m->d_ws, p.workspace_bytes,
                                   m->d_info, s);
    }
    if (exact_first_bit >= 0) {
        return hd_decode_selfsync_shard(m->d_units + p.halo_units, 0, p.shard_units, p.readable_units - p.halo_units,
                                        (uint32_t)exact_first_bit, m->table, m->d_out, p.out_capacity, m->d_ws,
                                        p.workspace_bytes, m->d_info, s);
    }
    if (synchronous)
        return hd_decode_selfsync_shard(m->d_units, p.halo_units, p.shard_units, p.readable_units, 0, m->table, m->d_out,
                                        p.out_capacity, m->d_ws, p.workspace_bytes, m->d_info, s);
    uint32_t *flag = m->h_flags + m->next_flag;
    m->next_flag = (m->next_flag + 1) % kFlagRing;
    return hd_decode_selfsync_shard_enqueue(m->d_units, p.halo_units, p.shard_units, p.readable_units, 0, m->table,
                                            m->d_out, p.out_capacity, m->d_ws, p.workspace_bytes, m->d_info, flag, s);
}

hd_status hd_mgpu_enqueue(hd_mgpu *m, void *stream) {
    if (!m || !m->table) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    hd_status st = mg_decode(m, -1, false, s);
    if (st != HD_OK) return st;
    m->pending = true;
    return mg_exchange(m, s);
}

hd_status hd_mgpu_align(hd_mgpu *m, void *stream) {
    if (!m) return HD_ERR_INVALID;
    return mg_exchange(m, static_cast<cudaStream_t>(stream));
}

static hd_status mg_read_infos(hd_mgpu *m, cudaStream_t s) {
    MG_CUDA(cudaMemcpyAsync(m->infos.data(), m->d_all, m->infos.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    MG_CUDA(cudaStreamSynchronize(s));
    if (m->mailbox && hd_mailbox_timed_out(m->mailbox)) return HD_ERR_UNRESOLVED;
    return HD_OK;
}

// Host half (collective): every decision is taken from the exchanged triples, which all ranks hold identically, so the
// ranks walk through the same sequence of repair rounds without any further agreement protocol.
hd_status hd_mgpu_finish(hd_mgpu *m, uint64_t n_symbols_total, uint64_t *offset, uint64_t *count, void *stream) {
    if (!m || !m->table) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!m->pending) {
        hd_status st = hd_mgpu_enqueue(m, stream);
        if (st != HD_OK) return st;
    }
    m->pending = false;
    m->repairs = 0;
    hd_status st = mg_read_infos(m, s);
    if (st != HD_OK) return st;
    for (uint32_t round = 0; round < 2 * m->world + 2; ++round) {
        // (1) shards whose synchronisation had not converged when they were decoded: decode again, to the fixed point
        bool any_unconverged = false;
        for (uint32_t g = 0; g < m->world; ++g)
            if (m->infos[3 * g + 1] != kEmpty && (m->infos[3 * g + 1] & kUnconvergedBit)) any_unconverged = true;
        if (any_unconverged) {
            if (m->infos[3 * m->rank + 1] != kEmpty && (m->infos[3 * m->rank + 1] & kUnconvergedBit)) {
                st = mg_decode(m, -1, true, s);
                if (st != HD_OK) return st;
                m->repairs += 1;
            }
        } else {
            // (2) seams: a rank whose halo was too short decodes again from the exact offset its predecessor left at
            const uint32_t bad = hd_mgpu_first_bad_seam(m->infos.data(), m->world);
            if (bad >= m->world) {
                hd_mgpu_place(m->infos.data(), m->world, n_symbols_total, m->rank, offset, count);
                return HD_OK;
            }
            if (bad == m->rank) {
                uint64_t prev_exit = 0;
                for (uint32_t g = 0; g < bad; ++g)
                    if (m->infos[3 * g + 1] != kEmpty) prev_exit = m->infos[3 * g + 2];
                st = mg_decode(m, (int)(prev_exit & 15u), true, s);
                if (st != HD_OK) return st;
                m->repairs += 1;
            }
        }
        st = mg_exchange(m, s);
        if (st != HD_OK) return st;
        st = mg_read_infos(m, s);
        if (st != HD_OK) return st;
    }
    return HD_ERR_UNRESOLVED;
}

uint32_t hd_mgpu_repair_count(const hd_mgpu *m) { return m ? m->repairs : 0; }

// End-to-end step from / to HOST memory: this rank's shard goes through the three-stage pipeline of the host decoder
// (pieces upload, decode and download on three streams: both PCIe directions busy at once), then the triple is
// exchanged like after hd_mgpu_enqueue.  The context is re-bound to the host decoder's device buffers, so that the rare
// seam repair of hd_mgpu_finish can run on the device; if it did (hd_mgpu_repair_count() != 0) the caller refreshes its
// host output with hd_mgpu_host_refresh.
hd_status hd_mgpu_host_step(hd_mgpu *m, hd_host_decoder *hdec, const uint32_t *h_units, const uint32_t *h_gap_words,
                            uint8_t *h_out, uint64_t out_cap, void *stream) {
    if (!m || !hdec || !m->table) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const hd_mgpu_plan &p = m->plan;
    uint64_t info[3] = {0, kEmpty, kEmpty};
    if (p.shard_units) {
        if (!h_units || !h_out) return HD_ERR_INVALID;
        hd_status st;
        if (m->scheme == HD_SCHEME_GAP) {
            if (!h_gap_words) return HD_ERR_INVALID;
            const bool from_prev = p.seg_begin != 0;
            st = hd_host_decode_gap_shard(hdec, h_units, p.readable_units, h_gap_words + (from_prev ? 1 : 0),
                                          p.seg_end - p.seg_begin, from_prev ? h_gap_words[0] >> 28 : 0u, m->table, h_out,
                                          out_cap, info, nullptr);
        } else {
            st = hd_host_decode_selfsync_shard(hdec, h_units, p.halo_units, p.shard_units, p.readable_units, m->table,
                                               h_out, out_cap, info, nullptr);
        }
        if (st != HD_OK) return st;
        uint32_t *du = nullptr, *dg = nullptr;
        uint8_t *dout = nullptr;
        void *dws = nullptr;
        hd_host_decoder_buffers(hdec, &du, &dg, &dout, &dws, nullptr);
        m->d_units = du;
        m->d_out = dout;
        m->d_ws = dws;
    }
    std::memcpy(m->h_info, info, sizeof(info));
    MG_CUDA(cudaMemcpyAsync(m->d_info, m->h_info, sizeof(info), cudaMemcpyHostToDevice, s));
    m->pending = true;
    return mg_exchange(m, s);
}

hd_status hd_mgpu_host_refresh(hd_mgpu *m, uint8_t *h_out, uint64_t count, void *stream) {
    if (!m || !h_out) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (count) MG_CUDA(cudaMemcpyAsync(h_out, m->d_out, count, cudaMemcpyDeviceToHost, s));
    MG_CUDA(cudaStreamSynchronize(s));
    return HD_OK;
}

hd_status hd_mgpu_infos(const hd_mgpu *m, uint64_t *infos_out) {
    if (!m || !infos_out) return HD_ERR_INVALID;
    std::memcpy(infos_out, m->infos.data(), m->infos.size() * sizeof(uint64_t));
    return HD_OK;
}



}  // extern "C"
