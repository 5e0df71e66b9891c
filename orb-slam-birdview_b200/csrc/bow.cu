// DBoW2 vocabulary transform (Frame::ComputeBoW, reference src/Frame.cc:562-569 ->
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1203, 1230-1271): every descriptor descends the vocabulary
// tree choosing, level by level, the child with the smallest Hamming distance (first minimum).  The descent (k*L
// distances per descriptor) runs on the device, one thread per descriptor; the ordered accumulation into
// BowVector / FeatureVector (std::map semantics, double sums in feature order) is done by the host side of the ABI.
#include "ctx.cuh"

namespace orbb200 {

struct VocDev {
    int nNodes, L;
    const int32_t* childPtr;
    const int32_t* childIdx;
    const uint4* desc;          // [nNodes][2]
};

__global__ void __launch_bounds__(128) bow_descend_kernel(VocDev V, const uint4* __restrict__ features, const int32_t* __restrict__ n_ptr, int n,
                                                          int levelsup, int32_t* __restrict__ outLeaf, int32_t* __restrict__ outNode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = n_ptr ? min(*n_ptr, n) : n;
    if (i >= cnt) return;
    const uint4 fa = __ldg(features + 2 * (size_t)i), fb = __ldg(features + 2 * (size_t)i + 1);
    const int nid_level = V.L - levelsup;
    int nid = 0, final_id = 0, level = 0;
    do {
        ++level;
        const int cb = __ldg(V.childPtr + final_id), ce = __ldg(V.childPtr + final_id + 1);
        int best = 0x7fffffff, bestId = final_id;
        for (int c = cb; c < ce; c++) {
            const int id = __ldg(V.childIdx + c);
            const uint4 da = __ldg(V.desc + 2 * (size_t)id), db = __ldg(V.desc + 2 * (size_t)id + 1);
            const int d = __popc(fa.x ^ da.x) + __popc(fa.y ^ da.y) + __popc(fa.z ^ da.z) + __popc(fa.w ^ da.w) +
                          __popc(fb.x ^ db.x) + __popc(fb.y ^ db.y) + __popc(fb.z ^ db.z) + __popc(fb.w ^ db.w);
            if (d < best) { best = d; bestId = id; }       // strict '<': first minimum in child order
        }
        final_id = bestId;
        if (level == nid_level) nid = final_id;
    } while (__ldg(V.childPtr + final_id) != __ldg(V.childPtr + final_id + 1) && level < 64);
    outLeaf[i] = final_id;
    outNode[i] = nid;
}

void launch_bow_descend(Ctx& c, const VocDev& V, const uint8_t* d_features, const int32_t* d_n, int n, int levelsup, int32_t* d_leaf, int32_t* d_node)
{
    if (n <= 0) return;
    bow_descend_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(V, reinterpret_cast<const uint4*>(d_features), d_n, n, levelsup, d_leaf, d_node);
    c.launches++;
}

}  // namespace orbb200

// ---- C ABI ------------------------------------------------------------------------------------------------
#include <algorithm>
#include <cmath>
#include <map>

using namespace orbb200;

struct orbb200_voc {
    Ctx* ctx = nullptr;
    VocDev dev{};
    void* owned[3] = {nullptr, nullptr, nullptr};
    std::vector<int32_t> wordId;        // per node, -1 for inner nodes
    std::vector<double> weight;         // per node
};

extern "C" {

int orbb200_voc_create(orbb200_ctx* ctx, orbb200_voc** out, int n_nodes, const int32_t* child_ptr, const int32_t* child_idx,
                       const uint8_t* node_desc, const int32_t* word_id, const double* weight, int L)
{
    if (!ctx) return ORBB200_ERR_ARG;
    Ctx& c = ctx->c;
    ORBB200_CUDA_OK(c, cudaSetDevice(c.device));
    if (!out || n_nodes <= 0 || !child_ptr || !child_idx || !node_desc || !word_id || !weight || L <= 0 || child_ptr[0] == child_ptr[1]) {
        c.err = "voc_create: bad argument"; return ORBB200_ERR_ARG;
    }
    orbb200_voc* v = new orbb200_voc;
    v->ctx = &c;
    const int nChild = child_ptr[n_nodes];
    bool ok = cudaMalloc(&v->owned[0], sizeof(int32_t) * (n_nodes + 1)) == cudaSuccess &&
              cudaMalloc(&v->owned[1], sizeof(int32_t) * std::max(nChild, 1)) == cudaSuccess &&
              cudaMalloc(&v->owned[2], (size_t)n_nodes * 32) == cudaSuccess;
    if (!ok) { c.err = "voc_create: cudaMalloc failed"; orbb200_voc_free(v); return ORBB200_ERR_CUDA; }
    cudaMemcpyAsync(v->owned[0], child_ptr, sizeof(int32_t) * (n_nodes + 1), cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(v->owned[1], child_idx, sizeof(int32_t) * nChild, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(v->owned[2], node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice, c.stream);
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    v->dev.nNodes = n_nodes; v->dev.L = L;
    v->dev.childPtr = (const int32_t*)v->owned[0]; v->dev.childIdx = (const int32_t*)v->owned[1]; v->dev.desc = (const uint4*)v->owned[2];
    v->wordId.assign(word_id, word_id + n_nodes);
    v->weight.assign(weight, weight + n_nodes);
    *out = v;
    return ORBB200_OK;
}

void orbb200_voc_free(orbb200_voc* v)
{
    if (!v) return;
    if (v->ctx) { cudaSetDevice(v->ctx->device); cudaStreamSynchronize(v->ctx->stream); }
    for (void* p : v->owned) if (p) cudaFree(p);
    delete v;
}

static int bow_finish(Ctx& c, const orbb200_voc* voc, const int32_t* d_leaf, const int32_t* d_node, int n,
                      int32_t* out_word, int32_t* out_node, int32_t* bow_word, double* bow_value, int* n_words,
                      int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* n_fv)
{
    std::vector<int32_t> leaf(std::max(n, 1)), node(std::max(n, 1));
    if (n > 0) {
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(leaf.data(), d_leaf, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(node.data(), d_node, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
    }
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    // BowVector::addWeight / FeatureVector::addFeature in feature order, then BowVector::normalize(L1)
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned> > fv;
    for (int i = 0; i < n; i++) {
        const double w = voc->weight[leaf[i]];
        if (out_word) out_word[i] = voc->wordId[leaf[i]];
        if (out_node) out_node[i] = node[i];
        if (w > 0) { v[(unsigned)voc->wordId[leaf[i]]] += w; fv[(unsigned)node[i]].push_back((unsigned)i); }
    }
    double norm = 0.0;
    for (auto it = v.begin(); it != v.end(); ++it) norm += fabs(it->second);
    if (norm > 0.0) for (auto it = v.begin(); it != v.end(); ++it) it->second /= norm;
    int k = 0;
    for (auto it = v.begin(); it != v.end(); ++it, ++k) { bow_word[k] = (int32_t)it->first; bow_value[k] = it->second; }
    *n_words = k;
    int nn = 0, p = 0;
    for (auto it = fv.begin(); it != fv.end(); ++it, ++nn) {
        fv_node[nn] = (int32_t)it->first; fv_ptr[nn] = p;
        for (unsigned idx : it->second) fv_idx[p++] = (int32_t)idx;
    }
    fv_ptr[nn] = p;
    *n_fv = nn;
    return ORBB200_OK;
}

int orbb200_bow_transform(orbb200_ctx* ctx, const orbb200_voc* voc, const uint8_t* desc, int n, int levelsup,
                          int32_t* out_word, int32_t* out_node, int32_t* bow_word, double* bow_value, int* n_words,
                          int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* n_fv)
{
    if (!ctx) return ORBB200_ERR_ARG;
    Ctx& c = ctx->c;
    ORBB200_CUDA_OK(c, cudaSetDevice(c.device));
    if (!voc || n < 0 || (n > 0 && !desc) || !bow_word || !bow_value || !n_words || !fv_node || !fv_ptr || !fv_idx || !n_fv) {
        c.err = "bow_transform: bad argument"; return ORBB200_ERR_ARG;
    }
    const size_t need = (size_t)std::max(n, 1) * (32 + 8) + 1024;
    if (!ensure_scratch(c, need, 0)) return ORBB200_ERR_CUDA;
    uint8_t* d_desc = c.d_scratch;
    int32_t* d_leaf = reinterpret_cast<int32_t*>(c.d_scratch + (((size_t)std::max(n, 1) * 32 + 255) & ~(size_t)255));
    int32_t* d_node = d_leaf + std::max(n, 1);
    if (n > 0) ORBB200_CUDA_OK(c, cudaMemcpyAsync(d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream));
    launch_bow_descend(c, voc->dev, d_desc, nullptr, n, levelsup, d_leaf, d_node);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return bow_finish(c, voc, d_leaf, d_node, n, out_word, out_node, bow_word, bow_value, n_words, fv_node, fv_ptr, fv_idx, n_fv);
}

int orbb200_bow_transform_extracted(orbb200_ctx* ctx, const orbb200_voc* voc, int img_index, int levelsup,
                                    int32_t* out_word, int32_t* out_node, int32_t* bow_word, double* bow_value, int* n_words,
                                    int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* n_fv)
{
    if (!ctx) return ORBB200_ERR_ARG;
    Ctx& c = ctx->c;
    ORBB200_CUDA_OK(c, cudaSetDevice(c.device));
    if (!voc || !c.cur || img_index < 0 || img_index >= c.curN || !bow_word || !bow_value || !n_words || !fv_node || !fv_ptr || !fv_idx || !n_fv) {
        c.err = "bow_transform_extracted: bad argument"; return ORBB200_ERR_ARG;
    }
    const int kpi = c.cur->g.kpPerImg;
    if (!ensure_scratch(c, (size_t)kpi * 8 + 1024, 0)) return ORBB200_ERR_CUDA;
    int32_t* d_leaf = reinterpret_cast<int32_t*>(c.d_scratch);
    int32_t* d_node = d_leaf + kpi;
    int32_t n = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&n, c.d_counts + img_index, sizeof(n), cudaMemcpyDeviceToHost, c.stream));
    launch_bow_descend(c, voc->dev, c.d_desc + (size_t)img_index * kpi * 32, c.d_counts + img_index, kpi, levelsup, d_leaf, d_node);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return bow_finish(c, voc, d_leaf, d_node, std::min(n, kpi), out_word, out_node, bow_word, bow_value, n_words, fv_node, fv_ptr, fv_idx, n_fv);
}

}  // extern "C"
