// Host interface of the persistent decode-step kernel (pstep.cu / pstep.cuh), used by the Taming engine (gpt.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wmar_b200.h"

namespace wmar {

struct PstepState;

struct PstepWeights {            // borrowed fp32 device pointers, reference nn.Linear layout [out][in]
    const float *tok_emb, *pos_emb, *lnf_g, *lnf_b, *head;
    struct Layer { const float *ln1_g, *ln1_b, *wqkv, *bqkv, *wproj, *bproj, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2; };
    const Layer *layers;         // [n_layer] host array
};

// true when the model tiles for the persistent kernel (head_dim 64, d <= 1536, V % 16 == 0, block_size <= 1024, plan fits)
bool pstep_eligible(const wmar_gpt_config &cfg, int n_sms);
// plans, allocates and packs (re-tiles the weights into ring stages; ~ one extra copy of the model in HBM)
int pstep_create(const wmar_gpt_config &cfg, int n_sms, const PstepWeights &w, float *kcache, float *vcache, float *logits,
                 const int *d_step, const int64_t *d_seq, int seq_ld, PstepState **out);
void pstep_destroy(PstepState *s);
// clears the epoch flags; once per generation (the token-step counter, part of every epoch, restarts at 0)
int pstep_reset(PstepState *s, cudaStream_t stream);
// enqueues one token step (graph-capturable)
int pstep_enqueue(PstepState *s, int B, int *d_err, cudaStream_t stream);
// probe only: copies the globaltimer stamps of the last step ([G][512]); returns G
int pstep_trace(PstepState *s, unsigned long long *out, size_t cap);

}  // namespace wmar
