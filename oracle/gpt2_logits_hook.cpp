// oracle/gpt2_logits_hook.cpp — TEST INFRASTRUCTURE ONLY.
//
// Lets the tests see the LOGITS of the reference's unmodified examples/gpt-2/main-backend.cpp: the example's only consumer of the
// logits is gpt_sample_top_k_top_p (examples/common.cpp, called at main-backend.cpp:895).  For the "-dump" variants of the gpt-2
// binaries (oracle/Makefile) examples/common.cpp is compiled with -Dgpt_sample_top_k_top_p=gpt_sample_top_k_top_p_orig and this file
// supplies gpt_sample_top_k_top_p: it appends the n_vocab logits of every sampling step to $GPT2_LOGITS_DUMP (raw f32) and, when
// $GPT2_FORCE_TOKENS names a file of int32 token ids, returns those instead of sampling (teacher forcing: two backends are then
// compared on the SAME token trajectory, step by step).  main-backend.cpp itself is compiled unmodified.
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

gpt_vocab::id gpt_sample_top_k_top_p_orig(const gpt_vocab & vocab, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng);

gpt_vocab::id gpt_sample_top_k_top_p(const gpt_vocab & vocab, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng) {
    static int step = 0;
    static std::vector<int32_t> forced;
    static bool loaded = false;
    if (!loaded) {
        loaded = true;
        if (const char * f = getenv("GPT2_FORCE_TOKENS")) {
            if (FILE * fp = fopen(f, "rb")) {
                int32_t v;
                while (fread(&v, sizeof(v), 1, fp) == 1) forced.push_back(v);
                fclose(fp);
            }
        }
    }
    if (const char * f = getenv("GPT2_LOGITS_DUMP")) {
        if (FILE * fp = fopen(f, step == 0 ? "wb" : "ab")) {
            fwrite(logits, sizeof(float), vocab.id_to_token.size(), fp);
            fclose(fp);
        }
    }
    gpt_vocab::id id = gpt_sample_top_k_top_p_orig(vocab, logits, top_k, top_p, temp, rng);
    if ((size_t) step < forced.size()) id = forced[(size_t) step];
    ++step;
    return id;
}
