// oracle/ref_probe.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim over the UNMODIFIED reference's *public* API (include/ggml.h,
// ggml-backend.h, ggml-cpu.h), built into oracle/_ref/libggml_probe.so and driven from
// Python with ctypes.  It exists so that tests/ and bench.py's cpu_baseline /
// `--impl reference` leg can (a) obtain the reference's own bytes/floats for
// quantize / dequantize / vec_dot / MUL_MAT / MUL_MAT_ID, and (b) drive ANY registered
// ggml backend (the reference CPU backend, or our plug-in loaded with
// ggml_backend_load) through the reference's own graph + backend API, the way
// tests/test-backend-ops.cpp does (build graph → alloc_ctx_tensors → tensor_set →
// graph_compute → tensor_get).  Nothing here is on the product path.

#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include "ggml-cpu.h"
#include "gguf.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct backend_holder {
    ggml_backend_t      be   = nullptr;
    ggml_threadpool_t   pool = nullptr;
    ~backend_holder() {
        if (be)   ggml_backend_free(be);
        if (pool) ggml_threadpool_free(pool);
    }
};

// the fp16->fp32 lookup table (ggml_table_f32_f16) is only filled by the first ggml_init()
struct table_init {
    table_init() { ggml_init_params ip = { 1024, nullptr, true }; ggml_free(ggml_init(ip)); ggml_cpu_init(); }
} g_table_init;

bool open_backend(backend_holder & h, const char * dev_name, int n_threads) {
    ggml_backend_dev_t dev = ggml_backend_dev_by_name(dev_name);
    if (!dev) { fprintf(stderr, "probe: no device named %s\n", dev_name); return false; }
    h.be = ggml_backend_dev_init(dev, nullptr);
    if (!h.be) return false;
    if (ggml_backend_is_cpu(h.be)) {
        if (n_threads <= 0) n_threads = (int) std::thread::hardware_concurrency();
        // persistent pool: avoids the disposable per-graph threadpool (ggml-cpu.c:14093)
        ggml_threadpool_params tp = ggml_threadpool_params_default(n_threads);
        h.pool = ggml_threadpool_new(&tp);
        ggml_backend_cpu_set_n_threads(h.be, n_threads);
        ggml_backend_cpu_set_threadpool(h.be, h.pool);
    }
    return true;
}
} // namespace

extern "C" {

int probe_load_backend(const char * path) { return ggml_backend_load(path) != nullptr; }
int probe_dev_count(void) { return (int) ggml_backend_dev_count(); }
const char * probe_dev_name(int i) { return ggml_backend_dev_name(ggml_backend_dev_get(i)); }
const char * probe_dev_desc(int i) { return ggml_backend_dev_description(ggml_backend_dev_get(i)); }
int probe_hw_threads(void) { return (int) std::thread::hardware_concurrency(); }

size_t probe_row_size(int type, int64_t k) { return ggml_row_size((ggml_type) type, k); }
int64_t probe_blck_size(int type) { return ggml_blck_size((ggml_type) type); }
size_t probe_type_size(int type) { return ggml_type_size((ggml_type) type); }

// ggml_quantize_chunk (src/ggml.c:6410) with imatrix == NULL
size_t probe_quantize(int type, const float * src, void * dst, int64_t nrows, int64_t k) {
    // the lowest-bit i-quants refuse to quantize without an importance matrix (ggml_quantize_requires_imatrix): give them uniform weights
    std::vector<float> ones;
    if (ggml_quantize_requires_imatrix((ggml_type) type)) ones.assign((size_t) k, 1.0f);
    return ggml_quantize_chunk((ggml_type) type, src, dst, 0, nrows, k, ones.empty() ? nullptr : ones.data());
}
// type_traits[type].from_float_ref == quantize_row_*_ref (src/ggml-quants.c)
void probe_quantize_row_ref(int type, const float * src, void * dst, int64_t k) {
    ggml_get_type_traits((ggml_type) type)->from_float_ref(src, dst, k);
}
// type_traits[type].to_float == dequantize_row_* (src/ggml-quants.c)
void probe_dequantize(int type, const void * src, float * dst, int64_t n) {
    ggml_get_type_traits((ggml_type) type)->to_float(src, dst, n);
}
// type_traits_cpu[type].from_float: the SIMD activation quantizers (ggml-cpu-quants.c)
void probe_cpu_from_float(int type, const float * src, void * dst, int64_t k) {
    ggml_cpu_init();
    ggml_get_type_traits_cpu((ggml_type) type)->from_float(src, dst, k);
}
int probe_cpu_vec_dot_type(int type) { return (int) ggml_get_type_traits_cpu((ggml_type) type)->vec_dot_type; }
// type_traits_cpu[type].vec_dot on one (weight row, quantized activation row) pair
float probe_cpu_vec_dot(int type, int64_t k, const void * wrow, const void * yq) {
    ggml_cpu_init();
    float s = 0.0f;
    ggml_get_type_traits_cpu((ggml_type) type)->vec_dot((int) k, &s, 0, wrow, 0, yq, 0, 1);
    return s;
}

// Y[N,M] = MUL_MAT(W[type_a; K x M (x nb02 x nb03)], X[f32; K x N x b2 x b3]) on device `dev`.
// ne2a/ne3a: batch dims of W; ne2b/ne3b: batch dims of X (broadcast as ggml does).
// `repeat` identical nodes are put in ONE graph (as eval_perf does, test-backend-ops.cpp:657)
// and the graph is computed `iters` times; returns seconds per mul_mat node (compute only).
// If e2e != 0 each iteration also does tensor_set(X) before and tensor_get(Y) after, and
// those copies are inside the timed region.  Returns <0 on failure.
double probe_mul_mat(const char * dev, int type_a, const void * W, const float * X, float * Y,
                     int64_t M, int64_t N, int64_t K,
                     int64_t ne2a, int64_t ne3a, int64_t ne2b, int64_t ne3b,
                     int n_threads, int repeat, int iters, int warmup, int e2e) {
    backend_holder h;
    if (!open_backend(h, dev, n_threads)) return -1.0;

    ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(repeat + 8) + ggml_graph_overhead_custom(repeat + 8, false), nullptr, true };
    ggml_context * ctx = ggml_init(ip);
    ggml_tensor * a = ggml_new_tensor_4d(ctx, (ggml_type) type_a, K, M, ne2a, ne3a);
    ggml_tensor * b = ggml_new_tensor_4d(ctx, GGML_TYPE_F32, K, N, ne2b, ne3b);
    std::vector<ggml_tensor *> outs;
    ggml_cgraph * gf = ggml_new_graph_custom(ctx, repeat + 8, false);
    for (int r = 0; r < repeat; ++r) {
        ggml_tensor * c = ggml_mul_mat(ctx, a, b);
        outs.push_back(c);
        ggml_build_forward_expand(gf, c);
    }
    if (!ggml_backend_supports_op(h.be, outs[0])) { ggml_free(ctx); return -2.0; }
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, h.be);
    if (!buf) { ggml_free(ctx); return -3.0; }
    ggml_backend_tensor_set(a, W, 0, ggml_nbytes(a));
    ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));

    for (int i = 0; i < warmup; ++i) ggml_backend_graph_compute(h.be, gf);
    ggml_backend_synchronize(h.be);
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) {
        if (e2e) ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));
        ggml_backend_graph_compute(h.be, gf);
        if (e2e) ggml_backend_tensor_get(outs.back(), Y, 0, ggml_nbytes(outs.back()));
    }
    ggml_backend_synchronize(h.be);
    double t1 = now_s();
    ggml_backend_tensor_get(outs.back(), Y, 0, ggml_nbytes(outs.back()));
    ggml_backend_buffer_free(buf);
    ggml_free(ctx);
    return (t1 - t0) / ((double) (iters > 0 ? iters : 1) * repeat);
}

// The bench workload on the reference's CPU backend: `nw` DISTINCT weight tensors (each filled with the same W bytes, but separate
// memory, so a sweep streams nw x the matrix from DRAM like the GPU arm's 13-matrix sweep instead of re-reading one cache-resident
// matrix), one shared x, nw MUL_MAT nodes in one graph.  Returns seconds per mul_mat node.  e2e: tensor_set(x) before and
// tensor_get of every y after, inside the timed region.
double probe_mul_mat_sweep(const char * dev, int type_a, const void * W, const float * X, float * Y,
                           int64_t M, int64_t N, int64_t K, int nw, int n_threads, int iters, int warmup, int e2e) {
    backend_holder h;
    if (!open_backend(h, dev, n_threads)) return -1.0;
    ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(2 * nw + 8) + ggml_graph_overhead_custom(nw + 8, false), nullptr, true };
    ggml_context * ctx = ggml_init(ip);
    ggml_tensor * b = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, N);
    std::vector<ggml_tensor *> ws, outs;
    ggml_cgraph * gf = ggml_new_graph_custom(ctx, nw + 8, false);
    for (int r = 0; r < nw; ++r) {
        ggml_tensor * a = ggml_new_tensor_2d(ctx, (ggml_type) type_a, K, M);
        ggml_tensor * c = ggml_mul_mat(ctx, a, b);
        ws.push_back(a);
        outs.push_back(c);
        ggml_build_forward_expand(gf, c);
    }
    if (!ggml_backend_supports_op(h.be, outs[0])) { ggml_free(ctx); return -2.0; }
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, h.be);
    if (!buf) { ggml_free(ctx); return -3.0; }
    for (ggml_tensor * a : ws) ggml_backend_tensor_set(a, W, 0, ggml_nbytes(a));
    ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));
    for (int i = 0; i < warmup; ++i) ggml_backend_graph_compute(h.be, gf);
    ggml_backend_synchronize(h.be);
    // e2e == 1: synchronous public calls (tensor_set, graph_compute, one tensor_get per output);
    // e2e == 2: the asynchronous public calls a pipelined application uses (tensor_set_async, graph_compute_async, tensor_get_async
    //           per output into distinct host rows, one synchronize per iteration); identical to mode 1 on backends without async copies
    // results land in the device's pinned host buffer type when it offers one (what an application that cares about async copies uses)
    std::vector<float> host_vec;
    ggml_backend_buffer_t host_buf = nullptr;
    float * host_out_p = nullptr;
    if (e2e == 2) {
        ggml_backend_buffer_type_t hbt = ggml_backend_dev_host_buffer_type(ggml_backend_get_device(h.be));
        if (hbt) host_buf = ggml_backend_buft_alloc_buffer(hbt, (size_t) nw * M * N * sizeof(float) + K * N * sizeof(float));
        if (host_buf) host_out_p = (float *) ggml_backend_buffer_get_base(host_buf);
        else { host_vec.resize((size_t) nw * M * N + (size_t) K * N); host_out_p = host_vec.data(); }
        memcpy(host_out_p + (size_t) nw * M * N, X, (size_t) K * N * sizeof(float));      // x staged in the same (pinned) buffer
    }
    const float * host_x = e2e == 2 ? host_out_p + (size_t) nw * M * N : X;
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) {
        if (e2e == 2) {
            ggml_backend_tensor_set_async(h.be, b, host_x, 0, ggml_nbytes(b));
            ggml_backend_graph_compute_async(h.be, gf);
            for (int r = 0; r < nw; ++r) ggml_backend_tensor_get_async(h.be, outs[r], host_out_p + (size_t) r * M * N, 0, ggml_nbytes(outs[r]));
            ggml_backend_synchronize(h.be);
            continue;
        }
        if (e2e) ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));
        ggml_backend_graph_compute(h.be, gf);
        if (e2e) for (ggml_tensor * c : outs) ggml_backend_tensor_get(c, Y, 0, ggml_nbytes(c));
    }
    ggml_backend_synchronize(h.be);
    double t1 = now_s();
    ggml_backend_tensor_get(outs.back(), Y, 0, ggml_nbytes(outs.back()));
    if (host_buf) ggml_backend_buffer_free(host_buf);
    ggml_backend_buffer_free(buf);
    ggml_free(ctx);
    return (t1 - t0) / ((double) (iters > 0 ? iters : 1) * nw);
}

// MUL_MAT with the weights in the backend's SPLIT buffer type (rows sharded over all of the registry's devices), obtained the way an
// application does: ggml_backend_reg_get_proc_address(reg, "ggml_backend_split_buffer_type")(main_device, tensor_split)
// (include/ggml-backend.h:187-200; llama.cpp's usage).  Activations / result live in the main device's default buffer.
// tensor_split may be NULL (equal shares).  Returns seconds per graph_compute, < 0 on failure (-4: no split buffer type).
double probe_mul_mat_split(const char * dev, int main_device, const float * tensor_split, int type_a, const void * W, const float * X, float * Y,
                           int64_t M, int64_t N, int64_t K, int iters) {
    backend_holder h;
    if (!open_backend(h, dev, 0)) return -1.0;
    ggml_backend_reg_t reg = ggml_backend_dev_backend_reg(ggml_backend_get_device(h.be));
    typedef ggml_backend_buffer_type_t (*split_fn)(int, const float *);
    split_fn fn = (split_fn) ggml_backend_reg_get_proc_address(reg, "ggml_backend_split_buffer_type");
    if (!fn) return -4.0;
    ggml_backend_buffer_type_t sbuft = fn(main_device, tensor_split);
    if (!sbuft) return -4.0;
    ggml_init_params ipw = { ggml_tensor_overhead() * 4, nullptr, true };
    ggml_init_params ipc = { ggml_tensor_overhead() * 8 + ggml_graph_overhead(), nullptr, true };
    ggml_context * cw = ggml_init(ipw), * cc = ggml_init(ipc);
    ggml_tensor * a = ggml_new_tensor_2d(cw, (ggml_type) type_a, K, M);
    ggml_tensor * b = ggml_new_tensor_2d(cc, GGML_TYPE_F32, K, N);
    ggml_tensor * c = ggml_mul_mat(cc, a, b);
    ggml_cgraph * gf = ggml_new_graph(cc);
    ggml_build_forward_expand(gf, c);
    ggml_backend_buffer_t bw = ggml_backend_alloc_ctx_tensors_from_buft(cw, sbuft);
    if (!bw) { ggml_free(cw); ggml_free(cc); return -3.0; }
    ggml_backend_buffer_t bc = ggml_backend_alloc_ctx_tensors(cc, h.be);
    if (!bc) { ggml_backend_buffer_free(bw); ggml_free(cw); ggml_free(cc); return -3.0; }
    double out = -2.0;
    if (ggml_backend_supports_op(h.be, c)) {
        ggml_backend_tensor_set(a, W, 0, ggml_nbytes(a));
        ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));
        ggml_backend_graph_compute(h.be, gf);
        double t0 = now_s();
        for (int i = 0; i < iters; ++i) ggml_backend_graph_compute(h.be, gf);
        ggml_backend_synchronize(h.be);
        out = (now_s() - t0) / (iters > 0 ? iters : 1);
        ggml_backend_tensor_get(c, Y, 0, ggml_nbytes(c));
        // the split tensor must read back exactly as it was set (scatter / gather of the row ranges)
        std::vector<char> back(ggml_nbytes(a));
        ggml_backend_tensor_get(a, back.data(), 0, ggml_nbytes(a));
        if (memcmp(back.data(), W, ggml_nbytes(a)) != 0) out = -5.0;
    }
    ggml_backend_buffer_free(bc);
    ggml_backend_buffer_free(bw);
    ggml_free(cc); ggml_free(cw);
    return out;
}

// The on-disk format either side of the path: write W (quantized, given) as a GGUF file with the reference's own writer, read
// it back with the reference's own loader (gguf_init_from_file, no_alloc = true: metadata only), then do what a model loader does --
// allocate the tensors in the device's default buffer, stage the file's data section through the device's PINNED host buffer type
// (ggml_backend_dev_host_buffer_type) and upload with ggml_backend_tensor_set_async -- and run MUL_MAT on it.  Returns 0 and fills Y, or < 0.
double probe_gguf_mul_mat(const char * dev, const char * path, int type_a, const void * W, const float * X, float * Y, int64_t M, int64_t N, int64_t K) {
    {   // ---- write
        ggml_init_params ip = { ggml_tensor_overhead() * 4, nullptr, true };
        ggml_context * c = ggml_init(ip);
        ggml_tensor * a = ggml_new_tensor_2d(c, (ggml_type) type_a, K, M);
        ggml_set_name(a, "blk.0.ffn_up.weight");
        gguf_context * g = gguf_init_empty();
        gguf_set_val_str(g, "general.architecture", "probe");
        gguf_set_val_u32(g, "probe.rows", (uint32_t) M);
        gguf_add_tensor(g, a);
        gguf_set_tensor_data(g, "blk.0.ffn_up.weight", W);
        const bool ok = gguf_write_to_file(g, path, false);
        gguf_free(g); ggml_free(c);
        if (!ok) return -6.0;
    }
    backend_holder h;
    if (!open_backend(h, dev, 0)) return -1.0;
    ggml_context * cmeta = nullptr;
    gguf_init_params gp = { /* no_alloc = */ true, /* ctx = */ &cmeta };
    gguf_context * g = gguf_init_from_file(path, gp);
    if (!g || !cmeta) return -7.0;
    double out = -2.0;
    ggml_tensor * a = ggml_get_tensor(cmeta, "blk.0.ffn_up.weight");
    const int64_t tid = gguf_find_tensor(g, "blk.0.ffn_up.weight");
    if (a && tid >= 0 && gguf_get_val_u32(g, gguf_find_key(g, "probe.rows")) == (uint32_t) M && a->type == (ggml_type) type_a && a->ne[0] == K && a->ne[1] == M) {
        ggml_backend_buffer_t bw = ggml_backend_alloc_ctx_tensors(cmeta, h.be);          // weights in device memory
        ggml_init_params ipc = { ggml_tensor_overhead() * 8 + ggml_graph_overhead(), nullptr, true };
        ggml_context * cc = ggml_init(ipc);
        ggml_tensor * b = ggml_new_tensor_2d(cc, GGML_TYPE_F32, K, N);
        ggml_tensor * c = ggml_mul_mat(cc, a, b);
        ggml_cgraph * gf = ggml_new_graph(cc);
        ggml_build_forward_expand(gf, c);
        ggml_backend_buffer_t bc = ggml_backend_alloc_ctx_tensors(cc, h.be);
        if (bw && bc && ggml_backend_supports_op(h.be, c)) {
            // stage the tensor's bytes from the file through pinned host memory, upload asynchronously
            const size_t nbytes = ggml_nbytes(a), off = gguf_get_data_offset(g) + gguf_get_tensor_offset(g, tid);
            ggml_backend_buffer_type_t hbt = ggml_backend_dev_host_buffer_type(ggml_backend_get_device(h.be));
            ggml_backend_buffer_t hb = hbt ? ggml_backend_buft_alloc_buffer(hbt, nbytes) : nullptr;
            std::vector<char> pageable;
            void * stage = hb ? ggml_backend_buffer_get_base(hb) : (pageable.resize(nbytes), (void *) pageable.data());
            FILE * f = fopen(path, "rb");
            const bool rd = f && fseek(f, (long) off, SEEK_SET) == 0 && fread(stage, 1, nbytes, f) == nbytes;
            if (f) fclose(f);
            if (rd) {
                ggml_backend_tensor_set_async(h.be, a, stage, 0, nbytes);
                ggml_backend_tensor_set_async(h.be, b, X, 0, ggml_nbytes(b));
                ggml_backend_graph_compute_async(h.be, gf);
                ggml_backend_tensor_get_async(h.be, c, Y, 0, ggml_nbytes(c));
                ggml_backend_synchronize(h.be);
                out = hb ? 0.0 : 1.0;                              // 1: the device offers no pinned host buffer type (pageable staging)
            } else out = -8.0;
            if (hb) ggml_backend_buffer_free(hb);
        }
        if (bc) ggml_backend_buffer_free(bc);
        if (bw) ggml_backend_buffer_free(bw);
        ggml_free(cc);
    } else out = -9.0;
    gguf_free(g);
    ggml_free(cmeta);
    return out;
}

// C[M, n_used, n_tok] = MUL_MAT_ID(as[type; K x M x n_expert], b[f32; K x nb1 x n_tok], ids[i32; n_used x n_tok])
// (src/ggml.c:2735).  ids is given as the dense [n_ids_total x n_tok] array of which the
// first n_used columns are viewed (as test_mul_mat_id does, test-backend-ops.cpp:2017).
double probe_mul_mat_id(const char * dev, int type_a, const void * W, const float * X, const int32_t * ids, float * Y,
                        int64_t M, int64_t K, int64_t n_expert, int64_t n_used, int64_t nb1, int64_t n_tok,
                        int n_threads, int iters) {
    backend_holder h;
    if (!open_backend(h, dev, n_threads)) return -1.0;
    ggml_init_params ip = { ggml_tensor_overhead() * 16 + ggml_graph_overhead(), nullptr, true };
    ggml_context * ctx = ggml_init(ip);
    ggml_tensor * as  = ggml_new_tensor_3d(ctx, (ggml_type) type_a, K, M, n_expert);
    ggml_tensor * b   = ggml_new_tensor_3d(ctx, GGML_TYPE_F32, K, nb1, n_tok);
    ggml_tensor * idt = ggml_new_tensor_2d(ctx, GGML_TYPE_I32, n_used, n_tok);
    ggml_tensor * c   = ggml_mul_mat_id(ctx, as, b, idt);
    ggml_cgraph * gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, c);
    if (!ggml_backend_supports_op(h.be, c)) { ggml_free(ctx); return -2.0; }
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, h.be);
    if (!buf) { ggml_free(ctx); return -3.0; }
    ggml_backend_tensor_set(as, W, 0, ggml_nbytes(as));
    ggml_backend_tensor_set(b, X, 0, ggml_nbytes(b));
    ggml_backend_tensor_set(idt, ids, 0, ggml_nbytes(idt));
    ggml_backend_graph_compute(h.be, gf);
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) ggml_backend_graph_compute(h.be, gf);
    ggml_backend_synchronize(h.be);
    double t1 = now_s();
    ggml_backend_tensor_get(c, Y, 0, ggml_nbytes(c));
    ggml_backend_buffer_free(buf);
    ggml_free(ctx);
    return (t1 - t0) / (iters > 0 ? iters : 1);
}

} // extern "C"
