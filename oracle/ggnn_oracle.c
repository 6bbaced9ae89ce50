/* CPU ORACLE, plain C restatement -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * A second, independent statement of the reference's propagation step, written message by message in double precision:
 *   chem_tensorflow_sparse.py:117-218  (SparseGGNNChemModel.compute_final_node_representations, incl. the attention branch :170-196)
 *   chem_tensorflow_dense.py:93-117    (DenseGGNNChemModel.compute_final_node_representations)
 *   TF-1.3 rnn_cell_impl.py GRUCell / BasicRNNCell (un-vendored tensorflow==1.3.0, requirements.txt:2; restated from the release:
 *     [r|u] = sigmoid([x,h].K_g + b_g),  c = act([x, r*h].K_c + b_c),  h' = u*h + (1-u)*c ;   h' = act([x,h].K + b))
 *   tf.contrib.cudnn_rnn CudnnCompatibleGRUCell (sparse:105-108; TF >= 1.4): same gates, c = tanh(x.K_in + b_in + r*(h.K_hid + b_hid))
 * It shares no code with oracle/ggnn_oracle.py; tests/test_oracle.py requires the two to agree to 1e-12 on the golden fixtures and on
 * random batches, which is what pins each against transcription slips (PARITY UNPINNED BY THE REFERENCE itself: it ships no tests or
 * vectors and TensorFlow 1.3 cannot run here -- see the header of ggnn_oracle.py).
 *
 * Build (done by __graft_entry__.build()):  gcc -O2 -shared -fPIC -o oracle/libggnn_oracle_c.so oracle/ggnn_oracle.c -lm
 * All matrices are row-major doubles; index arrays are int32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMALL_NUMBER 1e-7 /* utils.py:8 */

typedef struct {
    int32_t hidden_size;                  /* D */
    int32_t num_edge_types;               /* T */
    int32_t num_layers;                   /* L */
    const int32_t* layer_timesteps;       /* [L]            sparse:131 */
    const int32_t* residual_offsets;      /* [L+1]          sparse:140-145 */
    const int32_t* residual_layers;       /* indices into node_states_per_layer */
    int32_t use_edge_bias;                /* sparse:202 */
    int32_t use_edge_msg_avg_aggregation; /* sparse:206 */
    int32_t use_propagation_attention;    /* sparse:170 */
    int32_t cell_is_rnn;                  /* 0 GRUCell, 1 BasicRNNCell, 2 CudnnCompatibleGRUCell   sparse:102-112 */
    int32_t act_is_relu;                  /* 0 tanh, 1 relu              sparse:75-81 */
} oracle_config;

typedef struct {
    const double* edge_weights;                /* [T][D][D] */
    const double* edge_biases;                 /* [T][D] or NULL */
    const double* edge_type_attention_weights; /* [T] or NULL */
    const double* gate_kernel;                 /* [(Din+D)][2D] (GRU) */
    const double* gate_bias;                   /* [2D] */
    const double* cand_kernel;                 /* [(Din+D)][D]  (RNN: the only kernel) */
    const double* cand_bias;                   /* [D] */
    const double* cand_hidden_bias;            /* [D] CudnnCompatibleGRUCell: candidate/hidden_projection/bias, else NULL */
} oracle_layer;

static double act(double v, int relu) { return relu ? (v > 0.0 ? v : 0.0) : tanh(v); }
static double sigmoid(double v) { return 1.0 / (1.0 + exp(-v)); }

/* out[n] = bias[n] + sum_k in[k] * K[k][n]   (the `_linear` of TF-1.3: concat(inputs) . kernel + bias) */
static void linear(const double* in, int K, const double* kernel, const double* bias, int N, double* out) {
    for (int n = 0; n < N; ++n) out[n] = bias[n];
    for (int k = 0; k < K; ++k) {
        const double a = in[k];
        if (a == 0.0) continue;
        const double* row = kernel + (size_t)k * N;
        for (int n = 0; n < N; ++n) out[n] += a * row[n];
    }
}

/* One cell call on one node: x [Din], h [D] -> hnew [D].  scratch: Din + 2D + 2D doubles. */
static void cell(const oracle_config* c, const oracle_layer* w, const double* x, int Din, const double* h, double* hnew, double* scratch) {
    const int D = c->hidden_size;
    double* in = scratch;                 /* [Din + D] */
    double* ru = scratch + Din + D;       /* [2D] */
    memcpy(in, x, sizeof(double) * (size_t)Din);
    if (c->cell_is_rnn == 1) {
        memcpy(in + Din, h, sizeof(double) * (size_t)D);
        linear(in, Din + D, w->cand_kernel, w->cand_bias, D, hnew);
        for (int d = 0; d < D; ++d) hnew[d] = act(hnew[d], c->act_is_relu);
        return;
    }
    memcpy(in + Din, h, sizeof(double) * (size_t)D);
    linear(in, Din + D, w->gate_kernel, w->gate_bias, 2 * D, ru);
    for (int d = 0; d < 2 * D; ++d) ru[d] = sigmoid(ru[d]);                 /* r = ru[0..D), u = ru[D..2D) */
    if (c->cell_is_rnn == 2) {
        /* tf.contrib.cudnn_rnn.CudnnCompatibleGRUCell (TF >= 1.4, sparse:105-108; restated from the release):
         *   c = tanh(_linear(x; input_projection) + r * _linear(h; hidden_projection)),  h' = u*h + (1-u)*c.
         * cand_kernel stacks the two kernels: rows [0, Din) input projection (bias cand_bias), rows [Din, Din+D) hidden projection. */
        double* hh = in + Din;            /* [D], reuses the concat buffer's tail */
        linear(h, D, w->cand_kernel + (size_t)Din * D, w->cand_hidden_bias, D, hh);
        linear(x, Din, w->cand_kernel, w->cand_bias, D, hnew);
        for (int d = 0; d < D; ++d) {
            const double cand = act(hnew[d] + ru[d] * hh[d], c->act_is_relu);
            hnew[d] = ru[D + d] * h[d] + (1.0 - ru[D + d]) * cand;
        }
        return;
    }
    for (int d = 0; d < D; ++d) in[Din + d] = ru[d] * h[d];                 /* [x, r*h] */
    linear(in, Din + D, w->cand_kernel, w->cand_bias, D, hnew);
    for (int d = 0; d < D; ++d) {
        const double cand = act(hnew[d], c->act_is_relu);
        hnew[d] = ru[D + d] * h[d] + (1.0 - ru[D + d]) * cand;
    }
}

/* sparse:117-218.  adjacency_lists[t] -> [num_edges[t]][2] (source, target); indeg [V][T]; h0, out [V][D].
 * Returns 0, or -1 on an out-of-range edge / allocation failure. */
int ggnn_oracle_sparse(const oracle_config* c, const oracle_layer* layers, int32_t V, const int32_t* const* adjacency_lists,
                       const int32_t* num_edges, const double* indeg, const double* h0, double* out) {
    const int D = c->hidden_size, T = c->num_edge_types, L = c->num_layers;
    const size_t VD = (size_t)V * D;
    int64_t M = 0;
    for (int t = 0; t < T; ++t) M += num_edges[t];
    double* states = (double*)malloc(sizeof(double) * VD * (size_t)(L + 1) + 8);   /* node_states_per_layer */
    double* incoming = (double*)malloc(sizeof(double) * (VD + 1));
    double* hnext = (double*)malloc(sizeof(double) * (VD + 1));
    double* score = (double*)malloc(sizeof(double) * (size_t)(M + 1));
    double* mx = (double*)malloc(sizeof(double) * (size_t)(V + 1));
    double* ssum = (double*)malloc(sizeof(double) * (size_t)(V + 1));
    const int din_max = D * (1 + 4);
    double* x = (double*)malloc(sizeof(double) * (size_t)din_max);
    double* scratch = (double*)malloc(sizeof(double) * (size_t)(din_max + 3 * D));
    double* msg = (double*)malloc(sizeof(double) * (size_t)D);
    if (!states || !incoming || !hnext || !score || !mx || !ssum || !x || !scratch || !msg) return -1;
    int rc = 0;
    memcpy(states, h0, sizeof(double) * VD);                                           /* sparse:118-119 */
    for (int l = 0; l < L && rc == 0; ++l) {
        const oracle_layer* w = &layers[l];
        const int nres = c->residual_offsets ? c->residual_offsets[l + 1] - c->residual_offsets[l] : 0;
        const int Din = D * (1 + nres);
        double* cur = states + (size_t)(l + 1) * VD;
        memcpy(cur, states + (size_t)l * VD, sizeof(double) * VD);                     /* sparse:152 */
        for (int s = 0; s < c->layer_timesteps[l] && rc == 0; ++s) {
            memset(incoming, 0, sizeof(double) * VD);
            if (c->use_propagation_attention) {                                        /* sparse:170-194, message by message */
                for (int v = 0; v < V; ++v) { mx[v] = -INFINITY; ssum[v] = 0.0; }
                int64_t m = 0;
                for (int t = 0; t < T; ++t)
                    for (int i = 0; i < num_edges[t]; ++i, ++m) {
                        const int src = adjacency_lists[t][2 * i], tgt = adjacency_lists[t][2 * i + 1];
                        if (src < 0 || src >= V || tgt < 0 || tgt >= V) { rc = -1; goto done; }
                        double dot = 0.0;
                        for (int d = 0; d < D; ++d) dot += cur[(size_t)src * D + d] * cur[(size_t)tgt * D + d];
                        score[m] = dot * w->edge_type_attention_weights[t];
                        if (score[m] > mx[tgt]) mx[tgt] = score[m];
                    }
                m = 0;
                for (int t = 0; t < T; ++t)
                    for (int i = 0; i < num_edges[t]; ++i, ++m) {
                        const int tgt = adjacency_lists[t][2 * i + 1];
                        score[m] = exp(score[m] - mx[tgt]);
                        ssum[tgt] += score[m];
                    }
            }
            int64_t m = 0;
            for (int t = 0; t < T; ++t) {                                              /* sparse:159-168,198 */
                const double* Wt = w->edge_weights + (size_t)t * D * D;
                for (int i = 0; i < num_edges[t]; ++i, ++m) {
                    const int src = adjacency_lists[t][2 * i], tgt = adjacency_lists[t][2 * i + 1];
                    if (src < 0 || src >= V || tgt < 0 || tgt >= V) { rc = -1; goto done; }
                    for (int d = 0; d < D; ++d) msg[d] = 0.0;
                    for (int k = 0; k < D; ++k) {
                        const double a = cur[(size_t)src * D + k];
                        for (int d = 0; d < D; ++d) msg[d] += a * Wt[(size_t)k * D + d];
                    }
                    const double alpha = c->use_propagation_attention ? score[m] / (ssum[tgt] + SMALL_NUMBER) : 1.0;
                    for (int d = 0; d < D; ++d) incoming[(size_t)tgt * D + d] += alpha * msg[d];
                }
            }
            for (int v = 0; v < V; ++v) {
                double deg = 0.0;
                for (int t = 0; t < T; ++t) {
                    const double n = indeg[(size_t)v * T + t];
                    deg += n;
                    if (c->use_edge_bias)                                              /* sparse:202-204 */
                        for (int d = 0; d < D; ++d) incoming[(size_t)v * D + d] += n * w->edge_biases[(size_t)t * D + d];
                }
                if (c->use_edge_msg_avg_aggregation)                                   /* sparse:206-209 */
                    for (int d = 0; d < D; ++d) incoming[(size_t)v * D + d] /= (deg + SMALL_NUMBER);
                for (int r = 0; r < nres; ++r)                                         /* sparse:211-212: residuals first, messages last */
                    memcpy(x + (size_t)r * D, states + (size_t)c->residual_layers[c->residual_offsets[l] + r] * VD + (size_t)v * D,
                           sizeof(double) * (size_t)D);
                memcpy(x + (size_t)nres * D, incoming + (size_t)v * D, sizeof(double) * (size_t)D);
                cell(c, w, x, Din, cur + (size_t)v * D, hnext + (size_t)v * D, scratch); /* sparse:215-216 */
            }
            memcpy(cur, hnext, sizeof(double) * VD);
        }
    }
    memcpy(out, states + (size_t)L * VD, sizeof(double) * VD);                         /* sparse:218 */
done:
    free(states); free(incoming); free(hnext); free(score); free(mx); free(ssum); free(x); free(scratch); free(msg);
    return rc;
}

/* dense:93-117.  adjacency [b][T][v][v] with A[g][t][dest][src]; h0, out [b*v][D]; one weight set, GRU/tanh, bias added to every row. */
int ggnn_oracle_dense(int32_t D, int32_t T, int32_t num_timesteps, int32_t use_edge_bias, const oracle_layer* w, int32_t b, int32_t v,
                      const double* adjacency, const double* h0, double* out) {
    const size_t V = (size_t)b * v, VD = V * D;
    oracle_config c;
    memset(&c, 0, sizeof c);
    c.hidden_size = D; c.num_edge_types = T;
    double* h = (double*)malloc(sizeof(double) * (VD + 1));
    double* hn = (double*)malloc(sizeof(double) * (VD + 1));
    double* m = (double*)malloc(sizeof(double) * (VD + 1));
    double* acts = (double*)malloc(sizeof(double) * (VD + 1));
    double* scratch = (double*)malloc(sizeof(double) * (size_t)(4 * D + 2 * D));
    if (!h || !hn || !m || !acts || !scratch) return -1;
    memcpy(h, h0, sizeof(double) * VD);
    for (int s = 0; s < num_timesteps; ++s) {
        memset(acts, 0, sizeof(double) * VD);
        for (int t = 0; t < T; ++t) {
            const double* Wt = w->edge_weights + (size_t)t * D * D;
            for (size_t r = 0; r < V; ++r) {                                           /* dense:103-108: m = h.W_t + b_t on EVERY row */
                double* mr = m + r * D;
                for (int d = 0; d < D; ++d) mr[d] = use_edge_bias ? w->edge_biases[(size_t)t * D + d] : 0.0;
                for (int k = 0; k < D; ++k) {
                    const double a = h[r * D + k];
                    for (int d = 0; d < D; ++d) mr[d] += a * Wt[(size_t)k * D + d];
                }
            }
            for (int g = 0; g < b; ++g)                                                /* dense:109-113: acts += A_t . m */
                for (int i = 0; i < v; ++i)
                    for (int j = 0; j < v; ++j) {
                        const double a = adjacency[(((size_t)g * T + t) * v + i) * v + j];
                        if (a != 0.0)
                            for (int d = 0; d < D; ++d) acts[((size_t)g * v + i) * D + d] += a * m[((size_t)g * v + j) * D + d];
                    }
        }
        for (size_t r = 0; r < V; ++r) cell(&c, w, acts + r * D, D, h + r * D, hn + r * D, scratch);   /* dense:115 */
        memcpy(h, hn, sizeof(double) * VD);
    }
    memcpy(out, h, sizeof(double) * VD);
    free(h); free(hn); free(m); free(acts); free(scratch);
    return 0;
}
