// oracle/gpt2_compare.cpp — TEST INFRASTRUCTURE ONLY.
//
// Node-by-node parity of the REAL gpt-2 token graph: the reference's examples/gpt-2/main-backend.cpp is included UNMODIFIED (its main renamed), its own
// gpt2_model_load / gpt2_graph build the model and the graph on the reference CPU backend, and the reference's own
// ggml_backend_compare_graph_backend (src/ggml-backend.cpp:1814: the machinery behind tests/test-backend-ops) evaluates every node on the
// CPU backend and on a second device (the B200 plug-in, loaded through $GGML_BACKEND_PATH), handing both results to a callback.  Prints one
// line per node: index, op, name, shape, NMSE(device, cpu); exit code 0.  Two graphs: the prompt batch (n_past = 0, N tokens) and one
// decode step (n_past = N, 1 token) whose KV cache was filled by the CPU evaluation of the prompt.
// With "sync" as 4th argument the device's copy of every node result is overwritten with the CPU's right after the comparison, so that every
// node is evaluated on IDENTICAL inputs on both sides: per-op parity on the real graph, free of error accumulation.  Without it the device
// graph runs on its own intermediate results (what a real run does) and the printed NMSE is the accumulated deviation.
// usage: gpt2-compare MODEL DEVICE [N_PROMPT] [sync]
#define main gpt2_example_main
#include "gpt-2/main-backend.cpp"
#undef main

#include <cinttypes>

namespace {
struct cmp_state { const char * tag; int n_bad; double worst; bool sync; int first_bad; char first_bad_op[64]; };

double nmse_f32(const float * a, const float * b, size_t n) {       // as tests/test-backend-ops.cpp:174-188 (a = device, b = cpu)
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < n; ++i) { const double d = (double) a[i] - (double) b[i]; num += d * d; den += (double) a[i] * (double) a[i]; }
    return den > 0.0 ? num / den : num;
}

bool on_node(int index, ggml_tensor * t1, ggml_tensor * t2, void * ud) {
    cmp_state * st = (cmp_state *) ud;
    if (t1->type != GGML_TYPE_F32 || !ggml_is_contiguous(t1)) return true;            // views / non-f32 nodes: compared through their consumers
    const size_t n = (size_t) ggml_nelements(t1);
    std::vector<float> a(n), b(n);
    ggml_backend_tensor_get(t1, b.data(), 0, n * sizeof(float));                    // t1: CPU (backend1)
    ggml_backend_tensor_get(t2, a.data(), 0, n * sizeof(float));                    // t2: device
    const double e = nmse_f32(a.data(), b.data(), n);
    if (e > st->worst) st->worst = e;
    if (e > 1e-9) { if (st->n_bad == 0) { st->first_bad = index; snprintf(st->first_bad_op, sizeof(st->first_bad_op), "%s", ggml_op_desc(t1)); } st->n_bad++; }
    if (st->sync) ggml_backend_tensor_set(t2, b.data(), 0, n * sizeof(float));     // the next nodes of the device graph consume the CPU's values
    printf("node %s %d %s %s [%" PRId64 ",%" PRId64 ",%" PRId64 ",%" PRId64 "] nmse %.3e\n", st->tag, index, ggml_op_desc(t1), t1->name, t1->ne[0], t1->ne[1], t1->ne[2], t1->ne[3], e);
    return true;
}
} // namespace

int main(int argc, char ** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s MODEL DEVICE [N_PROMPT]\n", argv[0]); return 2; }
    const int n_prompt = argc > 3 ? atoi(argv[3]) : 5;
    const bool sync = argc > 4 && strcmp(argv[4], "sync") == 0;
    ggml_backend_load_all();
    ggml_backend_dev_t dev = ggml_backend_dev_by_name(argv[2]);
    if (!dev) { fprintf(stderr, "no device %s\n", argv[2]); return 3; }
    ggml_backend_t be2 = ggml_backend_dev_init(dev, nullptr);
    gpt2_model model;
    gpt_vocab vocab;
    if (!gpt2_model_load(argv[1], model, vocab, 1024, 0)) return 4;                   // CPU backend (this TU is compiled without GGML_USE_CUDA)
    ggml_backend_cpu_set_n_threads(model.backend, 8);
    ggml_gallocr_t allocr = ggml_gallocr_new(ggml_backend_get_default_buffer_type(model.backend));
    std::vector<gpt_vocab::id> prompt;
    for (int i = 0; i < n_prompt; ++i) prompt.push_back(64 + (i * 7) % 60);
    int rc = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const int n_past = phase == 0 ? 0 : n_prompt, N = phase == 0 ? n_prompt : 1;
        ggml_cgraph * gf = gpt2_graph(model, n_past, N);
        ggml_gallocr_alloc_graph(allocr, gf);
        ggml_tensor * embd = ggml_graph_get_tensor(gf, "embd");
        std::vector<int32_t> toks(prompt.begin(), prompt.begin() + N);
        if (phase == 1) toks[0] = 99;
        ggml_backend_tensor_set(embd, toks.data(), 0, N * sizeof(int32_t));
        ggml_tensor * position = ggml_graph_get_tensor(gf, "position");
        for (int i = 0; i < N; ++i) { int32_t v = n_past + i; ggml_backend_tensor_set(position, &v, i * sizeof(int32_t), sizeof(v)); }
        cmp_state st{ phase == 0 ? "prompt" : "decode", 0, 0.0, sync, -1, "" };
        // evaluates every node on both backends (the CPU evaluation also fills the CPU-side KV cache that the decode phase copies over)
        if (!ggml_backend_compare_graph_backend(model.backend, be2, gf, on_node, &st)) { fprintf(stderr, "graph copy failed\n"); rc = 5; break; }
        printf("summary %s %s nodes_over_1e-9 %d worst %.3e first_over %d %s\n", st.tag, sync ? "sync" : "free", st.n_bad, st.worst, st.first_bad, st.first_bad_op[0] ? st.first_bad_op : "-");
    }
    ggml_gallocr_free(allocr);
    ggml_backend_free(be2);
    return rc;
}
