// Backbone engine: builds a layer plan per architecture and runs it natively (C++), so one
// orbit_engine_forward call enqueues the whole feature-extractor pass on the caller's stream.
//
// Reference: the timm model created in model/feature_extractors.py:37-79 and invoked at
// model/few_shot_recognisers.py:114-117,143-146; FiLM sites per model/film.py:38-74.
#include <atomic>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "convnet.cuh"
#include "gemm_tcgen05.cuh"
#include "vit.cuh"

namespace orbit {

struct ParamInfo {
    std::string name;
    int64_t numel, offset;
    int ndim = 1;
    int64_t dims[4] = {0, 0, 0, 0};
};

enum OpKind { OP_STEM, OP_DW, OP_SE, OP_PW, OP_SPATIAL_MEAN, OP_CONV3, OP_MAXPOOL, OP_PATCH, OP_ASSEMBLE, OP_LN, OP_ATTN };
enum Buf { BUF_X0 = 0, BUF_X1, BUF_E, BUF_D, BUF_H, BUF_PARTIAL, BUF_GATE, BUF_COL, BUF_COUNT, BUF_INPUT = 100, BUF_OUTPUT = 101, BUF_NONE = -1 };

struct Op {
    OpKind kind;
    int in = BUF_NONE, out = BUF_NONE, res = BUF_NONE;
    int cin = 0, cout = 0, k = 1, stride = 1, act = ACT_NONE;
    bool gated = false;         // PW: multiply A by the SE gate
    int64_t w = -1, b = -1;     // param offsets (conv weight / bias)
    int64_t w2 = -1, b2 = -1;   // SE expand
    int64_t fold = -1;          // derived offset of scale[C], shift[C]
    int fold_idx = -1;          // index into orbit_engine::folds
    int64_t dw_wt = -1;         // derived offset of re-laid-out depthwise weights
    int64_t w_split = -1;       // derived offset of the fp16 hi/lo split weights (PW, tcgen05 path)
    int64_t w_gemm = -1;        // CONV3: derived offset of the weights re-laid-out to [cout, kpad] (im2col k-order)
    int kpad = 0;               // CONV3: im2col row length (k*k*cin rounded up to a multiple of 4)
    bool same_pad = false;      // CONV3: TF SAME geometry (asymmetric for stride 2) instead of the symmetric `pad`
    int pad = 1;                // CONV3 / MAXPOOL: symmetric padding
    int save = BUF_NONE;        // CONV3: unused; PW-free resnet plumbing uses explicit buffers
    bool nchw_in = false;       // CONV3: input is the fp32 NCHW frame tensor
    bool bias_only = false;     // PW: a Linear layer: scale = 1, shift = bias (params + b)
    int ln = -1;                // LN: index into orbit_engine::lns
    bool cls_only = false;      // LN: normalise token 0 of every frame only and write [frames, dim] (final norm + token pooling)
    int heads = 0, patch = 0;   // ATTN / PATCH
    float eps = 0.f;
    int se_reduce = 0;
};

}  // namespace orbit

using namespace orbit;

struct orbit_engine {
    int arch = 0, feat_dim = 0;
    std::vector<ParamInfo> params, film;
    int64_t param_floats = 0, film_floats = 0, derived_floats = 0;
    int64_t ident = -1;  // derived offset of [ones(max_c) | zeros(max_c)] (identity scale/shift for calibration)
    int max_c = 0;
    std::vector<FoldEntry> folds;
    struct LnEntry { int64_t gamma, beta, film_gamma, film_beta, out; int dim; };
    std::vector<LnEntry> lns;   // LayerNorms: effective (FiLM-substituted) gamma/beta are copied to derived[out]
    int tokens = 0;             // ViT: tokens per frame (patches + class token)
    std::vector<Op> ops;
    int chunk_frames = 256;   // frames per pass through the layer plan (workspace ~10 MB per 224-px frame)
    int gemm_mode = 1;        // tcgen05 FP16x3
    int implicit_conv = 1;    // 3x3 stride-1 convolutions (Cin % 64 == 0) as implicit GEMM instead of im2col + GEMM
    int fuse_mbconv = 1;      // expand 1x1 -> depthwise in one kernel (mbx_kernel) where the block input has 16 / 24 channels:
                              // 0 = never; 1 = the stride-2 block with 16 input channels (B0 block 1.0: faster than the streaming
                              // expand GEMM + depthwise pair); 3 = every stride-2 block (also B0 block 2.0, which the unfused pair
                              // now matches); 2 = every supported block (block 1.1 too: measured slower)
    mutable std::atomic<int64_t> last_launches{0};
    // optional per-launch CUDA-event timing (option "profile"): one event before every launch + one at the end
    int profile = 0;
    struct ProfRec { int family; double bytes, flops; };
    mutable std::vector<cudaEvent_t> prof_events;
    mutable std::vector<ProfRec> prof_recs;
    mutable size_t prof_used = 0;
    mutable std::vector<size_t> prof_ends;   // index of the closing event of each forward call

    int64_t add_param(const std::string& name, int64_t d0, int64_t d1 = 0, int64_t d2 = 0, int64_t d3 = 0) {
        ParamInfo pi;
        pi.name = name;
        pi.dims[0] = d0; pi.dims[1] = d1; pi.dims[2] = d2; pi.dims[3] = d3;
        pi.ndim = d1 == 0 ? 1 : (d2 == 0 ? 2 : (d3 == 0 ? 3 : 4));
        pi.numel = d0 * (d1 ? d1 : 1) * (d2 ? d2 : 1) * (d3 ? d3 : 1);
        pi.offset = param_floats;
        params.push_back(pi);
        param_floats += (pi.numel + 3) / 4 * 4;  // every tensor starts 16-byte aligned (128-bit loads)
        return pi.offset;
    }
    int64_t add_derived(int64_t numel) {
        const int64_t o = derived_floats;
        derived_floats += (numel + 3) / 4 * 4;  // keep 16-byte alignment
        return o;
    }
    // BatchNorm: registers weight/bias/running_mean/running_var and a fold entry; returns derived offset
    int64_t add_bn(const std::string& name, int c, float eps, bool film_site, Op* op = nullptr) {
        FoldEntry f;
        if (op) op->fold_idx = (int)folds.size();
        max_c = std::max(max_c, c);
        f.gamma = add_param(name + ".weight", c);
        f.beta = add_param(name + ".bias", c);
        f.mean = add_param(name + ".running_mean", c);
        f.var = add_param(name + ".running_var", c);
        f.film_gamma = f.film_beta = -1;
        f.conv_bias = -1;
        f.channels = c;
        f.eps = eps;
        f.out = add_derived(2 * (int64_t)c);
        if (film_site) {
            film.push_back({name + ".weight", c, (int64_t)folds.size()});  // offset filled after sorting
            film.push_back({name + ".bias", c, (int64_t)folds.size()});
        }
        folds.push_back(f);
        return f.out;
    }
    // LayerNorm: registers weight/bias; FiLM site when `film_site` (model/film.py:57-66); returns index into lns
    int add_ln(const std::string& name, int dim, bool film_site) {
        LnEntry l;
        l.gamma = add_param(name + ".weight", dim);
        l.beta = add_param(name + ".bias", dim);
        l.film_gamma = l.film_beta = -1;
        l.out = add_derived(2 * (int64_t)dim);
        l.dim = dim;
        lns.push_back(l);
        if (film_site) {
            ParamInfo a; a.name = name + ".weight"; a.numel = dim; a.offset = -(int64_t)lns.size();
            ParamInfo b = a; b.name = name + ".bias";
            film.push_back(a); film.push_back(b);
        }
        return (int)lns.size() - 1;
    }
    void finalize_film() {
        // generator order = sorted names (feature_adapters.py:43-44); remember which fold entry each feeds
        std::sort(film.begin(), film.end(), [](const ParamInfo& a, const ParamInfo& b) { return a.name < b.name; });
        int64_t off = 0;
        for (auto& f : film) {
            const int64_t fold_idx = f.offset;
            f.offset = off;
            const bool is_weight = f.name.size() > 7 && f.name.compare(f.name.size() - 7, 7, ".weight") == 0;
            if (fold_idx >= 0) { if (is_weight) folds[fold_idx].film_gamma = off; else folds[fold_idx].film_beta = off; }
            else { LnEntry& l = lns[-fold_idx - 1]; if (is_weight) l.film_gamma = off; else l.film_beta = off; }
            off += f.numel;
        }
        film_floats = off;
    }
};

// One timm DepthwiseSeparableConv (expand == 1) / InvertedResidual block: [expand 1x1 + bn1 + SiLU] -> depthwise kxk +
// bn (FiLM site of InvertedResidual: bn2, model/film.py:39-40) + SiLU -> squeeze-excite (reduce = cin/4) -> gated
// project 1x1 + bn (+ skip). `*cur` is the buffer holding the block input; it is advanced to the block output.
static void add_mbconv(orbit_engine* e, const std::string& p, int cin, int cout, int k, int stride, int expand, float eps, int* cur_io) {
    const int cur = *cur_io;
    const int mid = cin * expand;
    const bool ds = expand == 1;
    int dw_in = cur;
    if (!ds) {  // expand 1x1 + bn1 + SiLU
        Op op; op.kind = OP_PW; op.in = cur; op.out = BUF_E; op.cin = cin; op.cout = mid; op.act = ACT_SILU;
        op.w = e->add_param(p + "conv_pw.weight", mid, cin, 1, 1);
        op.fold = e->add_bn(p + "bn1", mid, eps, false, &op);
        op.w_split = e->add_derived(2 * (int64_t)mid * cin);
        e->ops.push_back(op);
        dw_in = BUF_E;
    }
    {   // depthwise + bn + SiLU
        Op op; op.kind = OP_DW; op.in = dw_in; op.out = BUF_D; op.cin = op.cout = mid; op.k = k; op.stride = stride; op.act = ACT_SILU;
        op.w = e->add_param(p + "conv_dw.weight", mid, 1, k, k);
        op.fold = e->add_bn(p + (ds ? "bn1" : "bn2"), mid, eps, !ds, &op);
        op.dw_wt = e->add_derived((int64_t)mid * k * k);
        e->ops.push_back(op);
    }
    {   // squeeze-excite gate
        Op op; op.kind = OP_SE; op.in = BUF_PARTIAL; op.out = BUF_GATE; op.cin = op.cout = mid;
        op.se_reduce = std::max(1, cin / 4);
        op.w = e->add_param(p + "se.conv_reduce.weight", op.se_reduce, mid, 1, 1);
        op.b = e->add_param(p + "se.conv_reduce.bias", op.se_reduce);
        op.w2 = e->add_param(p + "se.conv_expand.weight", mid, op.se_reduce, 1, 1);
        op.b2 = e->add_param(p + "se.conv_expand.bias", mid);
        op.dw_wt = e->add_derived((int64_t)mid * op.se_reduce);   // expand weight transposed to [reduce][mid]
        e->ops.push_back(op);
    }
    {   // project 1x1 (gated input) + bn (+ residual)
        Op op; op.kind = OP_PW; op.in = BUF_D; op.cin = mid; op.cout = cout; op.act = ACT_NONE; op.gated = true;
        op.out = cur == BUF_X0 ? BUF_X1 : BUF_X0;
        if (stride == 1 && cin == cout) op.res = cur;
        op.w = e->add_param(p + (ds ? "conv_pw.weight" : "conv_pwl.weight"), cout, mid, 1, 1);
        op.fold = e->add_bn(p + (ds ? "bn2" : "bn3"), cout, eps, false, &op);
        op.w_split = e->add_derived(2 * (int64_t)cout * mid);
        e->ops.push_back(op);
        *cur_io = op.out;
    }
}

// ------------------------------------------------------------------------------------------------
// EfficientNet-B0 plan (timm tf_efficientnet_b0: TF SAME padding, BN eps 1e-3, SiLU, SE reduce = cin/4)
// ------------------------------------------------------------------------------------------------
static void build_efficientnet_b0(orbit_engine* e) {
    const float eps = 1e-3f;
    struct Stage { int repeats, k, stride, cout, expand; };
    const Stage stages[7] = {{1, 3, 1, 16, 1}, {2, 3, 2, 24, 6}, {2, 5, 2, 40, 6}, {3, 3, 2, 80, 6},
                             {3, 5, 1, 112, 6}, {4, 5, 2, 192, 6}, {1, 3, 1, 320, 6}};
    e->feat_dim = 1280;
    {
        Op op; op.kind = OP_STEM; op.in = BUF_INPUT; op.out = BUF_X0; op.cin = 3; op.cout = 32; op.k = 3; op.stride = 2; op.act = ACT_SILU;
        op.w = e->add_param("conv_stem.weight", 32, 3, 3, 3);
        op.fold = e->add_bn("bn1", 32, eps, true, &op);
        e->ops.push_back(op);
    }
    int cin = 32, cur = BUF_X0;
    for (int s = 0; s < 7; ++s) {
        for (int j = 0; j < stages[s].repeats; ++j) {
            const std::string p = "blocks." + std::to_string(s) + "." + std::to_string(j) + ".";
            add_mbconv(e, p, cin, stages[s].cout, stages[s].k, j == 0 ? stages[s].stride : 1, stages[s].expand, eps, &cur);
            cin = stages[s].cout;
        }
    }
    {   // conv_head + bn2 (FiLM, root) + SiLU, then global average pool
        Op op; op.kind = OP_PW; op.in = cur; op.out = BUF_H; op.cin = cin; op.cout = 1280; op.act = ACT_SILU;
        op.w = e->add_param("conv_head.weight", 1280, cin, 1, 1);
        op.fold = e->add_bn("bn2", 1280, eps, true, &op);
        op.w_split = e->add_derived(2 * (int64_t)1280 * cin);
        e->ops.push_back(op);
        Op pool; pool.kind = OP_SPATIAL_MEAN; pool.in = BUF_H; pool.out = BUF_OUTPUT; pool.cin = pool.cout = 1280;
        e->ops.push_back(pool);
    }
    e->finalize_film();
    e->ident = e->add_derived(2 * (int64_t)e->max_c);
}

// ------------------------------------------------------------------------------------------------
// EfficientNet-V2-S plan (timm tf_efficientnetv2_s: TF SAME padding, BN eps 1e-3, SiLU; reference
// model/feature_extractors.py:16-19). Stages: cn r2 k3 s1 c24 | er r4 k3 s2 e4 c48 | er r4 k3 s2 e4 c64 |
// ir r6 k3 s2 e4 c128 se | ir r9 k3 s1 e6 c160 se | ir r15 k3 s2 e6 c256 se | head 1280. The full 3x3 convs (stem, ConvBnAct,
// EdgeResidual.conv_exp) run as im2col + the tcgen05 GEMM with BN/FiLM/SiLU/skip in its epilogue. FiLM sites
// (model/film.py:38-46): root bn1/bn2, ConvBnAct.bn1, EdgeResidual.bn1, InvertedResidual.bn2.
// ------------------------------------------------------------------------------------------------
static void build_efficientnet_v2_s(orbit_engine* e) {
    const float eps = 1e-3f;
    e->feat_dim = 1280;
    e->chunk_frames = 128;   // the 112x112x216 im2col matrix is 10.8 MB per frame
    auto conv = [&](const std::string& wname, const std::string& bnname, int in_buf, int out_buf, int cin, int cout, int k,
                    int stride, int act, int res, bool film_site, bool nchw) {
        Op op; op.kind = OP_CONV3; op.in = in_buf; op.out = out_buf; op.res = res; op.cin = cin; op.cout = cout; op.k = k;
        op.stride = stride; op.same_pad = true; op.act = act; op.nchw_in = nchw;
        op.kpad = (k * k * cin + 3) / 4 * 4;
        op.w = e->add_param(wname, cout, cin, k, k);
        op.fold = e->add_bn(bnname, cout, eps, film_site, &op);
        op.w_gemm = e->add_derived((int64_t)cout * op.kpad);
        op.w_split = e->add_derived(2 * (int64_t)cout * op.kpad);
        e->ops.push_back(op);
    };
    conv("conv_stem.weight", "bn1", BUF_INPUT, BUF_X0, 3, 24, 3, 2, ACT_SILU, BUF_NONE, true, true);
    int cin = 24, cur = BUF_X0;
    struct Stage { char type; int repeats, k, stride, cout, expand; };
    const Stage stages[6] = {{'c', 2, 3, 1, 24, 1}, {'e', 4, 3, 2, 48, 4}, {'e', 4, 3, 2, 64, 4},
                             {'i', 6, 3, 2, 128, 4}, {'i', 9, 3, 1, 160, 6}, {'i', 15, 3, 2, 256, 6}};
    for (int s = 0; s < 6; ++s) {
        for (int j = 0; j < stages[s].repeats; ++j) {
            const std::string p = "blocks." + std::to_string(s) + "." + std::to_string(j) + ".";
            const int stride = j == 0 ? stages[s].stride : 1, cout = stages[s].cout;
            const int other = cur == BUF_X0 ? BUF_X1 : BUF_X0;
            const bool skip = stride == 1 && cin == cout;
            if (stages[s].type == 'c') {          // ConvBnAct: silu(bn1(conv x)) + x
                conv(p + "conv.weight", p + "bn1", cur, other, cin, cout, 3, stride, ACT_SILU, skip ? cur : BUF_NONE, true, false);
                cur = other;
            } else if (stages[s].type == 'e') {   // EdgeResidual: bn2(conv_pwl(silu(bn1(conv_exp x)))) + x
                const int mid = cin * stages[s].expand;
                conv(p + "conv_exp.weight", p + "bn1", cur, BUF_E, cin, mid, 3, stride, ACT_SILU, BUF_NONE, true, false);
                Op op; op.kind = OP_PW; op.in = BUF_E; op.out = other; op.cin = mid; op.cout = cout; op.act = ACT_NONE;
                if (skip) op.res = cur;
                op.w = e->add_param(p + "conv_pwl.weight", cout, mid, 1, 1);
                op.fold = e->add_bn(p + "bn2", cout, eps, false, &op);
                op.w_split = e->add_derived(2 * (int64_t)cout * mid);
                e->ops.push_back(op);
                cur = other;
            } else {
                add_mbconv(e, p, cin, cout, stages[s].k, stride, stages[s].expand, eps, &cur);
            }
            cin = cout;
        }
    }
    {
        Op op; op.kind = OP_PW; op.in = cur; op.out = BUF_H; op.cin = cin; op.cout = 1280; op.act = ACT_SILU;
        op.w = e->add_param("conv_head.weight", 1280, cin, 1, 1);
        op.fold = e->add_bn("bn2", 1280, eps, true, &op);
        op.w_split = e->add_derived(2 * (int64_t)1280 * cin);
        e->ops.push_back(op);
        Op pool; pool.kind = OP_SPATIAL_MEAN; pool.in = BUF_H; pool.out = BUF_OUTPUT; pool.cin = pool.cout = 1280;
        e->ops.push_back(pool);
    }
    e->finalize_film();
    e->ident = e->add_derived(2 * (int64_t)e->max_c);
}

// ------------------------------------------------------------------------------------------------
// Set encoder plan (reference model/set_encoders.py:81-120 SimplePrePoolNet): 5 x [conv3x3(pad 1, bias) +
// BatchNorm2d(eps 1e-5) + ReLU + maxpool 2x2] + global average pool -> 64-d per frame. State-dict keys as in the
// reference: encoder.layer{i}.0.{weight,bias}, encoder.layer{i}.1.{weight,bias,running_mean,running_var}.
// ------------------------------------------------------------------------------------------------
static void build_set_encoder(orbit_engine* e) {
    e->feat_dim = 64;
    int cin = 3, cur = BUF_INPUT;
    for (int i = 1; i <= 5; ++i) {
        const std::string p = "encoder.layer" + std::to_string(i) + ".";
        Op op; op.kind = OP_CONV3; op.in = cur; op.out = BUF_E; op.cin = cin; op.cout = 64; op.k = 3; op.act = ACT_RELU;
        op.nchw_in = (i == 1);
        op.kpad = (9 * cin + 3) / 4 * 4;
        op.pad = 1;
        op.w = e->add_param(p + "0.weight", 64, cin, 3, 3);
        op.b = e->add_param(p + "0.bias", 64);
        op.fold = e->add_bn(p + "1", 64, 1e-5f, false, &op);
        e->folds[op.fold_idx].conv_bias = op.b;
        op.w_gemm = e->add_derived((int64_t)64 * op.kpad);
        op.w_split = e->add_derived(2 * (int64_t)64 * op.kpad);
        e->ops.push_back(op);
        Op pool; pool.kind = OP_MAXPOOL; pool.in = BUF_E; pool.out = (i % 2) ? BUF_X0 : BUF_X1; pool.cin = pool.cout = 64;
        pool.k = 2; pool.stride = 2; pool.pad = 0;
        e->ops.push_back(pool);
        cur = pool.out;
        cin = 64;
    }
    Op mean; mean.kind = OP_SPATIAL_MEAN; mean.in = cur; mean.out = BUF_OUTPUT; mean.cin = mean.cout = 64;
    e->ops.push_back(mean);
    e->ident = e->add_derived(2 * (int64_t)e->max_c);
}

// ------------------------------------------------------------------------------------------------
// ViT plan (timm 0.6.12 vit_{small,base}_patch32_224*, num_classes=0, token pooling). State-dict keys as timm:
// patch_embed.proj.*, cls_token, pos_embed, [norm_pre.*], blocks.i.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*,
// norm.*. FiLM sites: every LayerNorm called norm/norm1/norm2 (model/film.py:57-66).
// ------------------------------------------------------------------------------------------------
static void build_vit(orbit_engine* e, int dim, int depth, int heads, float eps, bool pre_norm) {
    const int P = 32, np = 49, T = np + 1;
    e->feat_dim = dim;
    e->tokens = T;
    auto linear = [&](const std::string& name, int in_buf, int out_buf, int cin, int cout, int act, int res) {
        Op op; op.kind = OP_PW; op.in = in_buf; op.out = out_buf; op.res = res; op.cin = cin; op.cout = cout; op.act = act;
        op.bias_only = true;
        op.w = e->add_param(name + ".weight", cout, cin);
        op.b = e->add_param(name + ".bias", cout);
        op.w_split = e->add_derived(2 * (int64_t)cout * cin);
        e->max_c = std::max(e->max_c, cout);
        e->ops.push_back(op);
    };
    {
        Op col; col.kind = OP_PATCH; col.in = BUF_INPUT; col.out = BUF_COL; col.patch = P; col.cin = 3; col.cout = 3 * P * P;
        e->ops.push_back(col);
        Op op; op.kind = OP_PW; op.in = BUF_COL; op.out = BUF_D; op.cin = 3 * P * P; op.cout = dim; op.act = ACT_NONE; op.bias_only = true;
        op.w = e->add_param("patch_embed.proj.weight", dim, 3, P, P);
        op.b = e->add_param("patch_embed.proj.bias", dim);
        op.w_split = e->add_derived(2 * (int64_t)dim * 3 * P * P);
        op.patch = P;   // marks "rows = frames * patches" for this GEMM
        e->max_c = std::max(e->max_c, dim);
        e->ops.push_back(op);
        Op as; as.kind = OP_ASSEMBLE; as.in = BUF_D; as.out = BUF_X0; as.cin = as.cout = dim;
        as.w = e->add_param("cls_token", 1, 1, dim);
        as.b = e->add_param("pos_embed", 1, T, dim);
        e->ops.push_back(as);
    }
    auto layernorm = [&](const std::string& name, int in_buf, int out_buf, bool film_site, bool cls_only) {
        Op op; op.kind = OP_LN; op.in = in_buf; op.out = out_buf; op.cin = op.cout = dim; op.eps = eps; op.cls_only = cls_only;
        op.ln = e->add_ln(name, dim, film_site);
        e->ops.push_back(op);
    };
    if (pre_norm) layernorm("norm_pre", BUF_X0, BUF_X0, false, false);
    for (int i = 0; i < depth; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        layernorm(p + "norm1", BUF_X0, BUF_X1, true, false);
        linear(p + "attn.qkv", BUF_X1, BUF_E, dim, 3 * dim, ACT_NONE, BUF_NONE);
        { Op at; at.kind = OP_ATTN; at.in = BUF_E; at.out = BUF_D; at.cin = at.cout = dim; at.heads = heads; e->ops.push_back(at); }
        linear(p + "attn.proj", BUF_D, BUF_X0, dim, dim, ACT_NONE, BUF_X0);       // x += proj(attn)   (in-place residual)
        layernorm(p + "norm2", BUF_X0, BUF_X1, true, false);
        linear(p + "mlp.fc1", BUF_X1, BUF_E, dim, 4 * dim, ACT_GELU, BUF_NONE);
        linear(p + "mlp.fc2", BUF_E, BUF_X0, 4 * dim, dim, ACT_NONE, BUF_X0);    // x += fc2(gelu(fc1))
    }
    layernorm("norm", BUF_X0, BUF_OUTPUT, true, true);
    e->finalize_film();
    e->ident = e->add_derived(2 * (int64_t)e->max_c);
}

// ------------------------------------------------------------------------------------------------
// ResNet-18 plan (torchvision resnet18 with fc = Identity; BASELINE.json extension, not in the reference at this
// commit -- SURVEY.md F6). Every conv = im2col + GEMM with the folded BatchNorm (eps 1e-5) in the epilogue;
// BasicBlock: relu(bn2(conv2(relu(bn1(conv1 x)))) + shortcut(x)). FiLM sites (extension, by the reference's mechanism
// of substituting BatchNorm affine parameters): bn1 / bn2 of every BasicBlock.
// ------------------------------------------------------------------------------------------------
static void build_resnet18(orbit_engine* e) {
    const float eps = 1e-5f;
    e->feat_dim = 512;
    auto conv = [&](const std::string& wname, const std::string& bnname, int in_buf, int out_buf, int cin, int cout, int k,
                    int stride, int pad, int act, int res, bool film_site, bool nchw) {
        Op op; op.kind = OP_CONV3; op.in = in_buf; op.out = out_buf; op.res = res; op.cin = cin; op.cout = cout; op.k = k;
        op.stride = stride; op.pad = pad; op.act = act; op.nchw_in = nchw;
        op.kpad = (k * k * cin + 3) / 4 * 4;
        op.w = e->add_param(wname, cout, cin, k, k);
        op.fold = e->add_bn(bnname, cout, eps, film_site, &op);
        op.w_gemm = e->add_derived((int64_t)cout * op.kpad);
        op.w_split = e->add_derived(2 * (int64_t)cout * op.kpad);
        e->ops.push_back(op);
    };
    conv("conv1.weight", "bn1", BUF_INPUT, BUF_E, 3, 64, 7, 2, 3, ACT_RELU, BUF_NONE, false, true);
    { Op pool; pool.kind = OP_MAXPOOL; pool.in = BUF_E; pool.out = BUF_X0; pool.cin = pool.cout = 64; pool.k = 3; pool.stride = 2; pool.pad = 1; e->ops.push_back(pool); }
    int cin = 64, cur = BUF_X0;
    const int widths[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) {
        for (int b = 0; b < 2; ++b) {
            const std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(b) + ".";
            const int cout = widths[l], stride = (l > 0 && b == 0) ? 2 : 1;
            const int other = cur == BUF_X0 ? BUF_X1 : BUF_X0;
            int shortcut = cur;
            if (stride != 1 || cin != cout) {   // first: it must see the block INPUT geometry (writing BUF_D does not advance h, w)
                conv(p + "downsample.0.weight", p + "downsample.1", cur, BUF_D, cin, cout, 1, stride, 0, ACT_NONE, BUF_NONE, false, false);
                shortcut = BUF_D;
            }
            conv(p + "conv1.weight", p + "bn1", cur, BUF_E, cin, cout, 3, stride, 1, ACT_RELU, BUF_NONE, true, false);
            conv(p + "conv2.weight", p + "bn2", BUF_E, other, cout, cout, 3, 1, 1, 16 + ACT_RELU, shortcut, true, false);
            // torchvision registers conv1, bn1, conv2, bn2, downsample: param ORDER differs here, names are what matter
            cur = other;
            cin = cout;
        }
    }
    Op mean; mean.kind = OP_SPATIAL_MEAN; mean.in = cur; mean.out = BUF_OUTPUT; mean.cin = mean.cout = 512;
    e->ops.push_back(mean);
    e->finalize_film();
    e->ident = e->add_derived(2 * (int64_t)e->max_c);
}

// TF "SAME" geometry: out = ceil(in/s), pad_before = total/2 (stride 1 => symmetric (k-1)/2)
static void same_geometry(int in, int k, int s, int* out, int* pad_before) {
    *out = (in + s - 1) / s;
    const int total = std::max((*out - 1) * s + k - in, 0);
    *pad_before = total / 2;
}

// output size and top/left padding of a full convolution: symmetric `pad` or TF SAME
static void conv_geometry(const Op& op, int h, int w, int* ho, int* wo, int* pt, int* pl) {
    if (op.same_pad) {
        same_geometry(h, op.k, op.stride, ho, pt);
        same_geometry(w, op.k, op.stride, wo, pl);
    } else {
        *ho = (h + 2 * op.pad - op.k) / op.stride + 1;
        *wo = (w + 2 * op.pad - op.k) / op.stride + 1;
        *pt = *pl = op.pad;
    }
}

struct BufSizes {
    int64_t per_frame[BUF_COUNT];
};

// dry-run of the plan for an HxW frame: maximum floats per frame each workspace buffer must hold
static int plan_buffers(const orbit_engine* e, int H, int W, BufSizes* bs) {
    for (int i = 0; i < BUF_COUNT; ++i) bs->per_frame[i] = 0;
    int h = H, w = W;
    auto need = [&](int buf, int64_t n) { if (buf >= 0 && buf < BUF_COUNT) bs->per_frame[buf] = std::max(bs->per_frame[buf], n); };
    for (const Op& op : e->ops) {
        switch (op.kind) {
            case OP_STEM:
            case OP_DW: {
                int ho, wo, p;
                same_geometry(h, op.k, op.stride, &ho, &p);
                same_geometry(w, op.k, op.stride, &wo, &p);
                h = ho; w = wo;
                if (h < 1 || w < 1) return ORBIT_ERR_UNSUPPORTED;
                need(op.out, (int64_t)h * w * op.cout);
                if (op.kind == OP_DW) {
                    need(BUF_PARTIAL, (int64_t)dw_partial_groups(op.cout, h, w, op.k, op.stride) * op.cout);
                    if (op.cout % 2 == 0 && (op.k == 3 || op.k == 5))     // the fused expand + depthwise kernel tiles differently
                        need(BUF_PARTIAL, (int64_t)std::max(mbx_partial_groups(16, op.cout, h, w, op.k, op.stride), mbx_partial_groups(24, op.cout, h, w, op.k, op.stride)) * op.cout);
                }
                break;
            }
            case OP_SE: need(BUF_GATE, op.cout); break;
            case OP_PW: need(op.out, (int64_t)(e->tokens ? e->tokens : h * w) * op.cout); break;
            case OP_CONV3: {
                int ho, wo, pt, pl;
                conv_geometry(op, h, w, &ho, &wo, &pt, &pl);
                if (ho < 1 || wo < 1) return ORBIT_ERR_UNSUPPORTED;
                need(BUF_COL, (int64_t)ho * wo * op.kpad);
                need(op.out, (int64_t)ho * wo * op.cout);
                if (op.out != BUF_D) { h = ho; w = wo; }     // the downsample branch does not advance the main path
                break;
            }
            case OP_MAXPOOL:
                h = (h + 2 * op.pad - op.k) / op.stride + 1; w = (w + 2 * op.pad - op.k) / op.stride + 1;
                if (h < 1 || w < 1) return ORBIT_ERR_UNSUPPORTED;
                need(op.out, (int64_t)h * w * op.cout);
                break;
            case OP_SPATIAL_MEAN: break;
            case OP_PATCH:
                if (H % op.patch || W % op.patch || (H / op.patch) * (W / op.patch) + 1 != e->tokens) return ORBIT_ERR_UNSUPPORTED;
                need(op.out, (int64_t)(e->tokens - 1) * op.cout);
                break;
            case OP_ASSEMBLE:
            case OP_LN:
            case OP_ATTN:
                need(op.out, (int64_t)e->tokens * op.cout);
                break;
        }
    }
    return ORBIT_OK;
}

extern "C" int orbit_engine_create(orbit_engine** out, int arch) {
    if (!out) return ORBIT_ERR_ARG;
    orbit_engine* e = new orbit_engine();
    e->arch = arch;
    switch (arch) {
        case ORBIT_ARCH_EFFICIENTNET_B0: build_efficientnet_b0(e); break;
        case ORBIT_ARCH_EFFICIENTNET_V2_S: build_efficientnet_v2_s(e); break;
        case ORBIT_ARCH_SET_ENCODER: build_set_encoder(e); break;
        case ORBIT_ARCH_RESNET18: build_resnet18(e); break;
        case ORBIT_ARCH_VIT_S_32: build_vit(e, 384, 12, 6, 1e-6f, false); break;
        case ORBIT_ARCH_VIT_B_32: build_vit(e, 768, 12, 12, 1e-6f, false); break;
        case ORBIT_ARCH_VIT_B_32_CLIP: build_vit(e, 768, 12, 12, 1e-5f, true); break;
        default: delete e; return ORBIT_ERR_UNSUPPORTED;
    }
    *out = e;
    return ORBIT_OK;
}

extern "C" void orbit_engine_destroy(orbit_engine* e) {
    if (!e) return;
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    delete e;
}
extern "C" int orbit_engine_feat_dim(const orbit_engine* e) { return e ? e->feat_dim : ORBIT_ERR_ARG; }
extern "C" int orbit_engine_num_params(const orbit_engine* e) { return e ? (int)e->params.size() : ORBIT_ERR_ARG; }
extern "C" int64_t orbit_engine_param_floats(const orbit_engine* e) { return e ? e->param_floats : ORBIT_ERR_ARG; }
extern "C" int orbit_engine_num_film(const orbit_engine* e) { return e ? (int)e->film.size() : ORBIT_ERR_ARG; }
extern "C" int64_t orbit_engine_film_floats(const orbit_engine* e) { return e ? e->film_floats : ORBIT_ERR_ARG; }
extern "C" int64_t orbit_engine_derived_floats(const orbit_engine* e) { return e ? e->derived_floats : ORBIT_ERR_ARG; }
extern "C" int64_t orbit_engine_last_launches(const orbit_engine* e) { return e ? e->last_launches.load() : ORBIT_ERR_ARG; }

static int info(const std::vector<ParamInfo>& v, int i, char* name, int cap, int64_t* numel, int64_t* offset) {
    if (i < 0 || i >= (int)v.size()) return ORBIT_ERR_ARG;
    if (name && cap > 0) { std::strncpy(name, v[i].name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (numel) *numel = v[i].numel;
    if (offset) *offset = v[i].offset;
    return ORBIT_OK;
}
extern "C" int orbit_engine_param_info(const orbit_engine* e, int i, char* name, int cap, int64_t* numel, int64_t* offset) {
    return e ? info(e->params, i, name, cap, numel, offset) : ORBIT_ERR_ARG;
}
extern "C" int orbit_engine_param_shape(const orbit_engine* e, int i, int* ndim, int64_t* dims4) {
    if (!e || !ndim || !dims4 || i < 0 || i >= (int)e->params.size()) return ORBIT_ERR_ARG;
    *ndim = e->params[i].ndim;
    for (int d = 0; d < 4; ++d) dims4[d] = e->params[i].dims[d];
    return ORBIT_OK;
}
extern "C" int orbit_engine_film_info(const orbit_engine* e, int i, char* name, int cap, int64_t* numel, int64_t* offset) {
    return e ? info(e->film, i, name, cap, numel, offset) : ORBIT_ERR_ARG;
}

extern "C" int orbit_engine_set_option(orbit_engine* e, const char* key, int value) {
    if (!e || !key) return ORBIT_ERR_ARG;
    if (!std::strcmp(key, "chunk_frames")) { if (value < 1 || value > 4096) return ORBIT_ERR_ARG; e->chunk_frames = value; return ORBIT_OK; }
    if (!std::strcmp(key, "gemm")) { if (value < 0 || value > 2) return ORBIT_ERR_ARG; e->gemm_mode = value; return ORBIT_OK; }
    if (!std::strcmp(key, "profile")) { e->profile = value != 0; return ORBIT_OK; }
    if (!std::strcmp(key, "fuse_mbconv")) { if (value < 0 || value > 3) return ORBIT_ERR_ARG; e->fuse_mbconv = value; return ORBIT_OK; }
    if (!std::strcmp(key, "implicit_conv")) { e->implicit_conv = value != 0; return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}
extern "C" int orbit_engine_get_option(const orbit_engine* e, const char* key, int* value) {
    if (!e || !key || !value) return ORBIT_ERR_ARG;
    if (!std::strcmp(key, "chunk_frames")) { *value = e->chunk_frames; return ORBIT_OK; }
    if (!std::strcmp(key, "gemm")) { *value = e->gemm_mode; return ORBIT_OK; }
    if (!std::strcmp(key, "profile")) { *value = e->profile; return ORBIT_OK; }
    if (!std::strcmp(key, "fuse_mbconv")) { *value = e->fuse_mbconv; return ORBIT_OK; }
    if (!std::strcmp(key, "implicit_conv")) { *value = e->implicit_conv; return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}

extern "C" int orbit_engine_prepare(const orbit_engine* e, const float* params, const float* film, float* derived, void* stream) {
    if (!e || !params || !derived) return ORBIT_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_bn_fold(e->folds.data(), (int)e->folds.size(), params, film, derived, st);
    if (rc) return rc;
    if (e->ident >= 0) {
        rc = launch_fill_identity(derived + e->ident, e->max_c, st);
        if (rc) return rc;
    }
    for (const orbit_engine::LnEntry& l : e->lns) {
        rc = launch_ln_affine((film && l.film_gamma >= 0) ? film + l.film_gamma : params + l.gamma,
                              (film && l.film_beta >= 0) ? film + l.film_beta : params + l.beta, derived + l.out, l.dim, st);
        if (rc) return rc;
    }
    for (const Op& op : e->ops) {
        if (op.kind == OP_DW) {
            rc = launch_dw_relayout(params + op.w, op.cin, op.k * op.k, derived + op.dw_wt, st);
            if (rc) return rc;
        } else if (op.kind == OP_SE) {
            rc = launch_dw_relayout(params + op.w2, op.cin, op.se_reduce, derived + op.dw_wt, st);
            if (rc) return rc;
        } else if (op.kind == OP_PW && op.w_split >= 0) {
            rc = launch_weight_split(params + op.w, op.cout, op.cin, derived + op.w_split, st);
            if (rc) return rc;
        } else if (op.kind == OP_CONV3) {
            rc = launch_conv_weight_relayout(params + op.w, derived + op.w_gemm, op.cout, op.cin, op.k * op.k, op.kpad, op.nchw_in, st);
            if (rc) return rc;
            rc = launch_weight_split(derived + op.w_gemm, op.cout, op.kpad, derived + op.w_split, st);
            if (rc) return rc;
        }
    }
    return ORBIT_OK;
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int64_t orbit_engine_workspace_bytes(const orbit_engine* e, int height, int width) {
    if (!e || height <= 0 || width <= 0) return ORBIT_ERR_ARG;
    BufSizes bs;
    if (plan_buffers(e, height, width, &bs)) return ORBIT_ERR_UNSUPPORTED;
    int64_t total = 0;
    for (int i = 0; i < BUF_COUNT; ++i) total += align_up(bs.per_frame[i] * e->chunk_frames * (int64_t)sizeof(float), 1024);
    return total + 1024;
}

extern "C" int64_t orbit_engine_macs(const orbit_engine* e, int height, int width) {
    if (!e || height <= 0 || width <= 0) return ORBIT_ERR_ARG;
    BufSizes bs;
    if (plan_buffers(e, height, width, &bs)) return ORBIT_ERR_UNSUPPORTED;
    int64_t macs = 0;
    int h = height, w = width;
    for (const Op& op : e->ops) {
        switch (op.kind) {
            case OP_STEM:
            case OP_DW: {
                int ho, wo, p;
                same_geometry(h, op.k, op.stride, &ho, &p);
                same_geometry(w, op.k, op.stride, &wo, &p);
                h = ho; w = wo;
                macs += (int64_t)h * w * op.cout * op.k * op.k * (op.kind == OP_STEM ? op.cin : 1);
                break;
            }
            case OP_SE: macs += (int64_t)h * w * op.cin + 2LL * op.cin * op.se_reduce; break;   // squeeze adds + 2 FCs
            case OP_PW: {
                const int64_t rows = e->tokens ? (op.patch ? e->tokens - 1 : e->tokens) : (int64_t)h * w;
                macs += rows * op.cin * op.cout;
                break;
            }
            case OP_CONV3: {
                int ho, wo, pt, pl;
                conv_geometry(op, h, w, &ho, &wo, &pt, &pl);
                macs += (int64_t)ho * wo * op.k * op.k * op.cin * op.cout;
                if (op.out != BUF_D) { h = ho; w = wo; }
                break;
            }
            case OP_MAXPOOL:
                h = (h + 2 * op.pad - op.k) / op.stride + 1; w = (w + 2 * op.pad - op.k) / op.stride + 1;
                break;
            case OP_SPATIAL_MEAN: macs += (int64_t)h * w * op.cin; break;
            case OP_ATTN: macs += 2LL * e->tokens * e->tokens * op.cin; break;   // QK^T and PV
            case OP_PATCH: case OP_ASSEMBLE: case OP_LN: break;
        }
    }
    return macs;
}

static int prof_mark(const orbit_engine* e, cudaStream_t st) {
    if (e->prof_used == e->prof_events.size()) {
        cudaEvent_t ev;
        ORBIT_CUDA(cudaEventCreate(&ev));
        e->prof_events.push_back(ev);
    }
    ORBIT_CUDA(cudaEventRecord(e->prof_events[e->prof_used++], st));
    return ORBIT_OK;
}

// Sums the event-timed launches recorded since the last read, per kernel family (OpKind; 5 = calibration
// statistics). Blocks until the recorded work has finished. Arrays have ORBIT_PROFILE_FAMILIES entries.
extern "C" int orbit_engine_profile_read(const orbit_engine* e, double* ms, int64_t* launches, double* bytes, double* flops) {
    if (!e || !ms || !launches || !bytes || !flops) return ORBIT_ERR_ARG;
    for (int i = 0; i < ORBIT_PROFILE_FAMILIES; ++i) { ms[i] = 0; launches[i] = 0; bytes[i] = 0; flops[i] = 0; }
    if (e->prof_used == 0) return ORBIT_OK;
    ORBIT_CUDA(cudaEventSynchronize(e->prof_events[e->prof_used - 1]));
    // events: [k0 k1 ... kn-1 END] per forward call; records are in the same order without the END entries
    size_t ev = 0, end_i = 0;
    for (size_t r = 0; r < e->prof_recs.size(); ++r) {
        while (end_i < e->prof_ends.size() && ev == e->prof_ends[end_i]) { ++ev; ++end_i; }
        float t = 0.f;
        ORBIT_CUDA(cudaEventElapsedTime(&t, e->prof_events[ev], e->prof_events[ev + 1]));
        const auto& rec = e->prof_recs[r];
        ms[rec.family] += t; launches[rec.family] += 1; bytes[rec.family] += rec.bytes; flops[rec.family] += rec.flops;
        ++ev;
    }
    e->prof_recs.clear(); e->prof_ends.clear(); e->prof_used = 0;
    return ORBIT_OK;
}

// Runs the layer plan. With `calib` (mutable params blob) each normalised layer is run twice: once raw
// (identity scale/shift, no activation) to measure the per-channel batch statistics, which are written
// into running_mean/running_var (unbiased variance, as torch momentum=1 would) and folded, then normally.
static int run_plan(const orbit_engine* e, const float* params, float* calib, float* derived, const float* frames,
                    int num_frames, int height, int width, float* feats, void* workspace, int64_t workspace_bytes,
                    cudaStream_t st) {
    BufSizes bs;
    int rc = plan_buffers(e, height, width, &bs);
    if (rc) return rc;
    if (workspace_bytes < orbit_engine_workspace_bytes(e, height, width)) return ORBIT_ERR_WORKSPACE;
    if (calib && num_frames > e->chunk_frames) return ORBIT_ERR_UNSUPPORTED;  // statistics need ONE chunk

    float* buf[BUF_COUNT];
    {
        char* p = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 1024));
        for (int i = 0; i < BUF_COUNT; ++i) {
            buf[i] = reinterpret_cast<float*>(p);
            p += align_up(bs.per_frame[i] * e->chunk_frames * (int64_t)sizeof(float), 1024);
        }
    }
    int64_t launches = 0;
    for (int f0 = 0; f0 < num_frames; f0 += e->chunk_frames) {
        const int B = std::min(e->chunk_frames, num_frames - f0);
        const float* in_frames = frames + (int64_t)f0 * 3 * height * width;
        float* out_feats = feats + (int64_t)f0 * e->feat_dim;
        auto ptr = [&](int b) -> float* {
            if (b == BUF_INPUT) return const_cast<float*>(in_frames);
            if (b == BUF_OUTPUT) return out_feats;
            return b >= 0 ? buf[b] : nullptr;
        };
        int h = height, w = width, se_tiles = 0, se_hw = 0;
        for (size_t oi = 0; oi < e->ops.size(); ++oi) {
            const Op& op = e->ops[oi];
            const int passes = (calib && op.fold_idx >= 0) ? 2 : 1;
            int ho = h, wo = w;
            // MBConv front half in one kernel (expand 1x1 + bn1 + SiLU -> depthwise + bn2/FiLM + SiLU + SE partials): the
            // 6x-expanded tensor never reaches HBM. Not during BatchNorm calibration (needs the expand output's statistics).
            if (!calib && e->fuse_mbconv && e->gemm_mode != 0 && op.kind == OP_PW && !op.gated && op.act == ACT_SILU && op.res == BUF_NONE &&
                oi + 1 < e->ops.size() && e->ops[oi + 1].kind == OP_DW && e->ops[oi + 1].in == op.out &&
                e->ops[oi + 1].act == ACT_SILU && mbx_supported(op.cin, e->ops[oi + 1].k, e->ops[oi + 1].stride) && !e->tokens &&
                (e->fuse_mbconv == 2 || (e->ops[oi + 1].stride == 2 && (e->fuse_mbconv == 3 || op.cin == 16))) &&
                mbx_fits(op.cin, op.cout, (h + e->ops[oi + 1].stride - 1) / e->ops[oi + 1].stride,
                         (w + e->ops[oi + 1].stride - 1) / e->ops[oi + 1].stride, e->ops[oi + 1].k, e->ops[oi + 1].stride)) {
                const Op& dw = e->ops[oi + 1];
                int pt, pl;
                same_geometry(h, dw.k, dw.stride, &ho, &pt);
                same_geometry(w, dw.k, dw.stride, &wo, &pl);
                if (e->profile) { rc = prof_mark(e, st); if (rc) return rc; }
                rc = launch_mbconv_expand_dw(ptr(op.in), params + op.w, derived + op.fold, derived + op.fold + op.cout,
                                             derived + dw.dw_wt, derived + dw.fold, derived + dw.fold + dw.cout, ptr(dw.out),
                                             buf[BUF_PARTIAL], B, h, w, op.cin, op.cout, ho, wo, dw.k, dw.stride, pt, pl, st);
                if (rc) return rc;
                ++launches;
                se_tiles = mbx_partial_groups(op.cin, op.cout, ho, wo, dw.k, dw.stride); se_hw = ho * wo;
                if (e->profile)
                    e->prof_recs.push_back({(int)OP_DW, 4.0 * B * ((double)h * w * op.cin + (double)ho * wo * dw.cout + (double)se_tiles * dw.cout),
                                            2.0 * B * ((double)h * w * op.cin * op.cout + (double)dw.k * dw.k * dw.cout * ho * wo)});
                h = ho; w = wo;
                ++oi;            // the depthwise op is done too
                continue;
            }
            for (int pass = 0; pass < passes; ++pass) {
                const bool raw = passes == 2 && pass == 0;
                const float* scale = (raw || op.bias_only) ? derived + e->ident : derived + op.fold;
                const float* shift = raw ? derived + e->ident + e->max_c : (op.bias_only ? params + op.b : derived + op.fold + op.cout);
                const int act = raw ? ACT_NONE : op.act;
                double p_bytes = 0, p_flops = 0;
                if (e->profile) { rc = prof_mark(e, st); if (rc) return rc; }
                switch (op.kind) {
                    case OP_STEM: {
                        int pt, pl;
                        same_geometry(h, op.k, op.stride, &ho, &pt);
                        same_geometry(w, op.k, op.stride, &wo, &pl);
                        rc = launch_stem(ptr(op.in), params + op.w, scale, shift, ptr(op.out), B, h, w, ho, wo, pt, pl,
                                         op.cout, act, st);
                        p_bytes = 4.0 * B * (3.0 * h * w + (double)ho * wo * op.cout);
                        p_flops = 2.0 * 27 * op.cout * (double)B * ho * wo;
                        break;
                    }
                    case OP_DW: {
                        int pt, pl;
                        same_geometry(h, op.k, op.stride, &ho, &pt);
                        same_geometry(w, op.k, op.stride, &wo, &pl);
                        rc = launch_depthwise(ptr(op.in), derived + op.dw_wt, scale, shift, ptr(op.out),
                                              raw ? nullptr : buf[BUF_PARTIAL], B, h, w, op.cin, ho, wo, op.k, op.stride, pt,
                                              pl, act, st);
                        se_tiles = dw_partial_groups(op.cin, ho, wo, op.k, op.stride); se_hw = ho * wo;
                        p_bytes = 4.0 * B * op.cin * ((double)h * w + (double)ho * wo + se_tiles);
                        p_flops = 2.0 * op.k * op.k * op.cin * (double)B * ho * wo;
                        break;
                    }
                    case OP_SE:
                        rc = launch_se_gate(buf[BUF_PARTIAL], se_tiles, se_hw, params + op.w, params + op.b, derived + op.dw_wt,
                                            params + op.b2, buf[BUF_GATE], B, op.cin, op.se_reduce, st);
                        p_bytes = 4.0 * (B * op.cin * (se_tiles + 1.0) + 2.0 * op.cin * op.se_reduce);
                        p_flops = 4.0 * B * op.cin * op.se_reduce;
                        break;
                    case OP_PATCH:
                        rc = launch_patch_im2col(ptr(op.in), ptr(op.out), B, h, w, op.patch, st);
                        p_bytes = 8.0 * B * 3.0 * h * w;
                        break;
                    case OP_ASSEMBLE:
                        rc = launch_assemble_tokens(ptr(op.in), params + op.w, params + op.b, ptr(op.out), B, e->tokens - 1, op.cout, st);
                        p_bytes = 8.0 * B * e->tokens * op.cout;
                        break;
                    case OP_LN: {
                        const orbit_engine::LnEntry& l = e->lns[op.ln];
                        if (op.cls_only)
                            rc = launch_layernorm(ptr(op.in), (int64_t)e->tokens * op.cin, derived + l.out, derived + l.out + l.dim, op.eps,
                                                  ptr(op.out), op.cout, B, op.cin, st);
                        else
                            rc = launch_layernorm(ptr(op.in), op.cin, derived + l.out, derived + l.out + l.dim, op.eps, ptr(op.out),
                                                  op.cout, B * e->tokens, op.cin, st);
                        p_bytes = 8.0 * B * (op.cls_only ? 1 : e->tokens) * op.cin;
                        break;
                    }
                    case OP_ATTN:
                        rc = launch_attention(ptr(op.in), ptr(op.out), B, e->tokens, op.heads, op.cin / op.heads, st);
                        p_bytes = 16.0 * B * e->tokens * op.cin;
                        p_flops = 4.0 * B * e->tokens * e->tokens * op.cin;
                        break;
                    case OP_PW: {
                        const int M = e->tokens ? B * (op.patch ? e->tokens - 1 : e->tokens) : B * h * w;
                        const float* gate = op.gated ? buf[BUF_GATE] : nullptr;
                        const float* res = raw ? nullptr : ptr(op.res);
                        if (e->gemm_mode == 0 || raw) {
                            rc = launch_pointwise_ffma(ptr(op.in), params + op.w, scale, shift, gate, res, ptr(op.out), M,
                                                       op.cout, op.cin, h * w, act, st);
                        } else {
                            rc = launch_pointwise_tcgen05(ptr(op.in), derived + op.w_split, scale, shift, gate, res,
                                                          ptr(op.out), M, op.cout, op.cin, h * w, act,
                                                          e->gemm_mode == 1 ? 3 : 1, st);
                        }
                        p_bytes = 4.0 * ((double)M * op.cin + (double)M * op.cout * (res ? 2 : 1) + (double)op.cin * op.cout +
                                         (gate ? (double)B * op.cin : 0.0));
                        p_flops = 2.0 * M * (double)op.cin * op.cout;
                        break;
                    }
                    case OP_CONV3: {
                        int cho, cwo, cpt, cpl;
                        conv_geometry(op, h, w, &cho, &cwo, &cpt, &cpl);
                        const int M = B * cho * cwo;
                        // 3x3 stride-1 convolutions over >= 64 channels: implicit GEMM (no im2col matrix in HBM)
                        if (e->implicit_conv && e->gemm_mode == 1 && !raw && op.k == 3 && op.stride == 1 && cpt == 1 && cpl == 1 && !op.nchw_in &&
                            op.cin % 64 == 0 && (act == ACT_RELU || act == 16 + ACT_RELU || act == ACT_NONE || act == ACT_SILU)) {
                            const float* res_i = ptr(op.res);
                            rc = launch_conv3x3_tcgen05(ptr(op.in), derived + op.w_split, scale, shift, res_i, ptr(op.out), B, h, w, op.cin,
                                                        op.cout, act, st);
                            if (rc != ORBIT_ERR_UNSUPPORTED) {
                                p_bytes = 4.0 * ((double)M * op.cin + (double)M * op.cout * (res_i ? 2 : 1) + (double)op.kpad * op.cout);
                                p_flops = 2.0 * M * (double)op.kpad * op.cout;
                                if (op.out != BUF_D) { ho = cho; wo = cwo; }
                                break;
                            }
                        }
                        // first convolution on the 3-channel NCHW frames (resnet conv1, set-encoder layer 1): direct tensor-core conv
                        if (e->implicit_conv && e->gemm_mode == 1 && !raw && op.nchw_in && cpt == cpl && op.res == BUF_NONE) {
                            rc = launch_conv_first(ptr(op.in), params + op.w, scale, shift, ptr(op.out), B, h, w, op.cin, op.cout, op.k,
                                                   op.stride, cpt, cho, cwo, act, st);
                            if (rc != ORBIT_ERR_UNSUPPORTED) {
                                p_bytes = 4.0 * ((double)B * op.cin * h * w + (double)M * op.cout);
                                p_flops = 2.0 * M * (double)op.k * op.k * op.cin * op.cout;
                                if (op.out != BUF_D) { ho = cho; wo = cwo; }
                                break;
                            }
                        }
                        rc = launch_im2col(ptr(op.in), buf[BUF_COL], B, h, w, op.cin, op.k, op.stride, cpt, cpl, cho, cwo, op.kpad,
                                           op.nchw_in, st);
                        if (rc) return rc;
                        ++launches;
                        if (e->profile) { e->prof_recs.push_back({(int)OP_STEM, 4.0 * M * (op.cin + (double)op.kpad), 0.0}); rc = prof_mark(e, st); if (rc) return rc; }
                        const float* res = raw ? nullptr : ptr(op.res);
                        if (e->gemm_mode == 0 || raw)
                            rc = launch_pointwise_ffma(buf[BUF_COL], derived + op.w_gemm, scale, shift, nullptr, res, ptr(op.out),
                                                       M, op.cout, op.kpad, cho * cwo, act, st);
                        else
                            rc = launch_pointwise_tcgen05(buf[BUF_COL], derived + op.w_split, scale, shift, nullptr, res,
                                                          ptr(op.out), M, op.cout, op.kpad, cho * cwo, act, e->gemm_mode == 1 ? 3 : 1, st);
                        p_bytes = 4.0 * ((double)M * op.kpad + (double)M * op.cout * (res ? 2 : 1) + (double)op.kpad * op.cout);
                        p_flops = 2.0 * M * (double)op.kpad * op.cout;
                        if (op.out != BUF_D) { ho = cho; wo = cwo; }    // the downsample branch keeps the main path's geometry
                        break;
                    }
                    case OP_MAXPOOL:
                        ho = (h + 2 * op.pad - op.k) / op.stride + 1; wo = (w + 2 * op.pad - op.k) / op.stride + 1;
                        rc = launch_maxpool(ptr(op.in), ptr(op.out), B, h, w, op.cin, op.k, op.stride, op.pad, ho, wo, st);
                        p_bytes = 4.0 * B * op.cin * ((double)h * w + (double)ho * wo);
                        break;
                    case OP_SPATIAL_MEAN:
                        rc = launch_spatial_mean(ptr(op.in), ptr(op.out), B, h * w, op.cin, st);
                        p_bytes = 4.0 * B * op.cin * (h * w + 1.0);
                        p_flops = (double)B * op.cin * h * w;
                        break;
                }
                if (rc) return rc;
                ++launches;
                if (e->profile) {
                    int fam = (int)op.kind;
                    if (op.kind == OP_CONV3) fam = OP_PW;
                    else if (op.kind == OP_MAXPOOL || op.kind == OP_LN || op.kind == OP_ASSEMBLE) fam = OP_SPATIAL_MEAN;
                    else if (op.kind == OP_PATCH) fam = OP_STEM;
                    else if (op.kind == OP_ATTN) fam = OP_SE;
                    e->prof_recs.push_back({fam, p_bytes, p_flops});
                }
                if (raw) {
                    const FoldEntry& fe = e->folds[op.fold_idx];
                    if (e->profile) { rc = prof_mark(e, st); if (rc) return rc; e->prof_recs.push_back({5, 0.0, 0.0}); }
                    rc = launch_channel_stats(ptr(op.out), (int64_t)B * ho * wo, op.cout, calib + fe.mean, calib + fe.var, st);
                    if (rc) return rc;
                    if (fe.conv_bias >= 0) {   // statistics were taken before the conv bias: BN sees conv + bias
                        rc = launch_add_vec(calib + fe.mean, calib + fe.conv_bias, op.cout, st);
                        if (rc) return rc;
                    }
                    rc = launch_bn_fold(&fe, 1, calib, nullptr, derived, st);
                    if (rc) return rc;
                    launches += 2;
                }
            }
            h = ho; w = wo;
        }
    }
    if (e->profile) { rc = prof_mark(e, st); if (rc) return rc; e->prof_ends.push_back(e->prof_used - 1); }
    e->last_launches.store(launches);
    return ORBIT_OK;
}

static int check_forward_args(const orbit_engine* e, const float* params, const float* derived, const float* frames,
                              int num_frames, int height, int width, const float* feats, const void* workspace) {
    if (!e || !params || !derived || !frames || !feats || !workspace) return ORBIT_ERR_ARG;
    if (num_frames < 0 || height <= 0 || width <= 0) return ORBIT_ERR_ARG;
    if (!aligned16(frames) || !aligned16(feats) || !aligned16(params) || !aligned16(derived)) return ORBIT_ERR_UNSUPPORTED;
    return ORBIT_OK;
}

extern "C" int orbit_engine_forward(const orbit_engine* e, const float* params, const float* derived, const float* frames,
                                    int num_frames, int height, int width, float* feats, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
    const int rc = check_forward_args(e, params, derived, frames, num_frames, height, width, feats, workspace);
    if (rc) return rc;
    return run_plan(e, params, nullptr, const_cast<float*>(derived), frames, num_frames, height, width, feats, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

extern "C" int orbit_engine_calibrate(const orbit_engine* e, float* params, float* derived, const float* frames,
                                      int num_frames, int height, int width, float* feats, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    const int rc = check_forward_args(e, params, derived, frames, num_frames, height, width, feats, workspace);
    if (rc) return rc;
    if (num_frames < 2) return ORBIT_ERR_ARG;
    return run_plan(e, params, params, derived, frames, num_frames, height, width, feats, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

// =================================================================================================
// Training through the frozen extractor (first slice: MBConv networks = EfficientNet-B0).
// Reference: MultiStepFewShotRecogniser.personalise with adapt_features=True (model/few_shot_recognisers.py:196-198,
// 207-246): gradient steps on the FiLM parameters (affine weight/bias of the tagged BatchNorms, model/film.py:38-79)
// and a new linear head, BatchNorm in eval mode. forward_train keeps what the backward needs (raw conv outputs `c` of
// every BatchNorm'd layer that is followed by an activation or is a FiLM site, the activated depthwise outputs, the
// squeeze-excite means and gates); backward_train walks the plan in reverse. Gradients live in the workspace buffer with
// the same id as the activation they belong to.
// =================================================================================================
namespace {

struct TrainGeom {
    std::vector<int> in_h, in_w, out_h, out_w;
    std::vector<int64_t> c_off, a_off, m_off, g_off, wt_off;     // per-op offsets: saved arena (floats per FRAME), transposed weights
    int64_t saved_per_frame = 0, tderived_floats = 0;
    bool ok = true;
};

TrainGeom train_geometry(const orbit_engine* e, int H, int W) {
    TrainGeom t;
    const size_t n = e->ops.size();
    t.in_h.assign(n, 0); t.in_w.assign(n, 0); t.out_h.assign(n, 0); t.out_w.assign(n, 0);
    t.c_off.assign(n, -1); t.a_off.assign(n, -1); t.m_off.assign(n, -1); t.g_off.assign(n, -1); t.wt_off.assign(n, -1);
    if (e->tokens) { t.ok = false; return t; }
    int h = H, w = W;
    for (size_t i = 0; i < n; ++i) {
        const Op& op = e->ops[i];
        t.in_h[i] = h; t.in_w[i] = w;
        int ho = h, wo = w, p;
        switch (op.kind) {
            case OP_STEM: case OP_DW:
                same_geometry(h, op.k, op.stride, &ho, &p);
                same_geometry(w, op.k, op.stride, &wo, &p);
                t.c_off[i] = t.saved_per_frame; t.saved_per_frame += (int64_t)ho * wo * op.cout;
                if (op.kind == OP_DW) {
                    t.a_off[i] = t.saved_per_frame; t.saved_per_frame += (int64_t)ho * wo * op.cout;
                    t.m_off[i] = t.saved_per_frame; t.saved_per_frame += op.cout;
                }
                break;
            case OP_SE: t.g_off[i] = t.saved_per_frame; t.saved_per_frame += op.cout; break;
            case OP_PW:
                if (op.bias_only || op.patch) { t.ok = false; return t; }
                if (op.act == ACT_SILU) { t.c_off[i] = t.saved_per_frame; t.saved_per_frame += (int64_t)h * w * op.cout; }
                else if (op.act != ACT_NONE) { t.ok = false; return t; }
                t.wt_off[i] = t.tderived_floats; t.tderived_floats += 3 * (int64_t)op.cin * op.cout + 8;
                break;
            case OP_SPATIAL_MEAN: break;
            case OP_CONV3:       // set encoder: conv3x3 (pad 1, stride 1, bias) + BatchNorm + ReLU, followed by maxpool 2x2
                if (e->arch != ORBIT_ARCH_SET_ENCODER || op.k != 3 || op.stride != 1 || op.pad != 1 || op.act != ACT_RELU || op.cout != 64 ||
                    i + 1 >= n || e->ops[i + 1].kind != OP_MAXPOOL) { t.ok = false; return t; }
                t.c_off[i] = t.saved_per_frame; t.saved_per_frame += (int64_t)h * w * op.cout;
                if (!op.nchw_in) { t.wt_off[i] = t.tderived_floats; t.tderived_floats += 2 * (int64_t)op.cin * 9 * op.cout + 8; }
                break;
            case OP_MAXPOOL:
                if (e->arch != ORBIT_ARCH_SET_ENCODER || op.k != 2 || op.stride != 2 || op.pad != 0) { t.ok = false; return t; }
                ho = h / 2; wo = w / 2;
                if (ho < 1 || wo < 1) { t.ok = false; return t; }
                t.a_off[i] = t.saved_per_frame; t.saved_per_frame += (int64_t)ho * wo * op.cout;      // pooled output = next conv's input
                break;
            default: t.ok = false; return t;
        }
        t.out_h[i] = ho; t.out_w[i] = wo;
        h = ho; w = wo;
        t.saved_per_frame = (t.saved_per_frame + 3) / 4 * 4;
        t.tderived_floats = (t.tderived_floats + 3) / 4 * 4;
    }
    return t;
}

void workspace_buffers(const orbit_engine* e, const BufSizes& bs, void* workspace, float** buf) {
    char* p = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 1024));
    for (int i = 0; i < BUF_COUNT; ++i) {
        buf[i] = reinterpret_cast<float*>(p);
        p += align_up(bs.per_frame[i] * e->chunk_frames * (int64_t)sizeof(float), 1024);
    }
}

}  // namespace

extern "C" int64_t orbit_engine_train_saved_floats(const orbit_engine* e, int height, int width) {
    if (!e || height <= 0 || width <= 0) return ORBIT_ERR_ARG;
    const TrainGeom t = train_geometry(e, height, width);
    return t.ok ? t.saved_per_frame : (int64_t)ORBIT_ERR_UNSUPPORTED;
}

extern "C" int64_t orbit_engine_train_derived_floats(const orbit_engine* e) {
    if (!e) return ORBIT_ERR_ARG;
    const TrainGeom t = train_geometry(e, 224, 224);
    return t.ok ? t.tderived_floats : (int64_t)ORBIT_ERR_UNSUPPORTED;
}

// transposed (+ fp16 hi/lo split) 1x1 weights for the data-gradient GEMMs; the weights are frozen: once per model
extern "C" int orbit_engine_prepare_train(const orbit_engine* e, const float* params, float* tderived, void* stream) {
    if (!e || !params || !tderived) return ORBIT_ERR_ARG;
    const TrainGeom t = train_geometry(e, 224, 224);
    if (!t.ok) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const Op& op = e->ops[i];
        if (op.kind != OP_PW) continue;
        float* wt = tderived + t.wt_off[i];                       // [cin][cout]
        int rc = launch_transpose(params + op.w, wt, op.cout, op.cin, st);
        if (rc) return rc;
        rc = launch_weight_split(wt, op.cin, op.cout, wt + (int64_t)op.cin * op.cout, st);
        if (rc) return rc;
    }
    return ORBIT_OK;
}

extern "C" int orbit_engine_forward_train(const orbit_engine* e, const float* params, const float* derived, const float* frames,
                                          int num_frames, int height, int width, float* feats, float* saved, int64_t saved_floats,
                                          void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_forward_args(e, params, derived, frames, num_frames, height, width, feats, workspace);
    if (rc) return rc;
    if (!saved || !aligned16(saved)) return ORBIT_ERR_ARG;
    const TrainGeom t = train_geometry(e, height, width);
    if (!t.ok) return ORBIT_ERR_UNSUPPORTED;
    const int B = num_frames;
    if (B > e->chunk_frames) return ORBIT_ERR_UNSUPPORTED;        // one pass: the arena is indexed by frame
    if (saved_floats < t.saved_per_frame * B) return ORBIT_ERR_WORKSPACE;
    BufSizes bs;
    if ((rc = plan_buffers(e, height, width, &bs))) return rc;
    if (workspace_bytes < orbit_engine_workspace_bytes(e, height, width)) return ORBIT_ERR_WORKSPACE;
    float* buf[BUF_COUNT];
    workspace_buffers(e, bs, workspace, buf);
    cudaStream_t st = (cudaStream_t)stream;
    auto ptr = [&](int b) -> float* {
        if (b == BUF_INPUT) return const_cast<float*>(frames);
        if (b == BUF_OUTPUT) return feats;
        return b >= 0 ? buf[b] : nullptr;
    };
    const float* ones = derived + e->ident;
    const float* zeros = derived + e->ident + e->max_c;
    int64_t launches = 0;
    const float* gate = nullptr;
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const Op& op = e->ops[i];
        const int h = t.in_h[i], w = t.in_w[i], ho = t.out_h[i], wo = t.out_w[i];
        const float* scale = derived + op.fold;
        const float* shift = derived + op.fold + op.cout;
        float* c = t.c_off[i] >= 0 ? saved + t.c_off[i] * B : nullptr;
        switch (op.kind) {
            case OP_STEM: {
                int pt, pl, d;
                same_geometry(h, op.k, op.stride, &d, &pt);
                same_geometry(w, op.k, op.stride, &d, &pl);
                rc = launch_stem(ptr(op.in), params + op.w, ones, zeros, c, B, h, w, ho, wo, pt, pl, op.cout, ACT_NONE, st);
                if (rc) return rc;
                rc = launch_bn_act_forward(c, scale, shift, ptr(op.out), (int64_t)B * ho * wo, op.cout, op.act, st);
                launches += 2;
                break;
            }
            case OP_DW: {
                int pt, pl, d;
                same_geometry(h, op.k, op.stride, &d, &pt);
                same_geometry(w, op.k, op.stride, &d, &pl);
                rc = launch_depthwise(ptr(op.in), derived + op.dw_wt, ones, zeros, c, nullptr, B, h, w, op.cin, ho, wo, op.k, op.stride,
                                      pt, pl, ACT_NONE, st);
                if (rc) return rc;
                float* a = saved + t.a_off[i] * B;
                rc = launch_bn_act_forward(c, scale, shift, a, (int64_t)B * ho * wo, op.cout, op.act, st);
                if (rc) return rc;
                rc = launch_spatial_mean(a, saved + t.m_off[i] * B, B, ho * wo, op.cout, st);
                if (rc) return rc;
                rc = cudaMemcpyAsync(ptr(op.out), a, sizeof(float) * (size_t)B * ho * wo * op.cout, cudaMemcpyDeviceToDevice, st);
                launches += 4;
                break;
            }
            case OP_SE: {
                float* g = saved + t.g_off[i] * B;
                // the depthwise op right before this one holds the squeeze means
                rc = launch_se_gate(saved + t.m_off[i - 1] * B, 1, 1, params + op.w, params + op.b, derived + op.dw_wt, params + op.b2, g,
                                    B, op.cin, op.se_reduce, st);
                gate = g;
                ++launches;
                break;
            }
            case OP_PW: {
                const int M = B * h * w;
                if (op.act == ACT_SILU) {
                    rc = launch_pointwise_tcgen05(ptr(op.in), derived + op.w_split, ones, zeros, nullptr, nullptr, c, M, op.cout, op.cin,
                                                  h * w, ACT_NONE, 3, st);
                    if (rc) return rc;
                    rc = launch_bn_act_forward(c, scale, shift, ptr(op.out), M, op.cout, op.act, st);
                    launches += 2;
                } else {
                    rc = launch_pointwise_tcgen05(ptr(op.in), derived + op.w_split, scale, shift, op.gated ? gate : nullptr, ptr(op.res),
                                                  ptr(op.out), M, op.cout, op.cin, h * w, ACT_NONE, 3, st);
                    ++launches;
                }
                break;
            }
            case OP_SPATIAL_MEAN: {
                // set encoder: the pooled output of the last block lives in the saved arena
                const float* in = (i > 0 && e->ops[i - 1].kind == OP_MAXPOOL) ? saved + t.a_off[i - 1] * B : ptr(op.in);
                rc = launch_spatial_mean(in, ptr(op.out), B, h * w, op.cin, st);
                ++launches;
                break;
            }
            case OP_CONV3: {
                // raw convolution (no bias: it is folded into shift) into the arena; im2col + the tcgen05 GEMM as in run_plan
                const float* in = op.nchw_in ? ptr(op.in) : saved + t.a_off[i - 1] * B;
                const int M = B * h * w;
                rc = ORBIT_ERR_UNSUPPORTED;
                if (e->implicit_conv && !op.nchw_in && op.cin % 64 == 0)      // layers 2-5: implicit GEMM, no im2col matrix
                    rc = launch_conv3x3_tcgen05(in, derived + op.w_split, ones, zeros, nullptr, c, B, h, w, op.cin, op.cout, ACT_NONE, st);
                else if (e->implicit_conv && op.nchw_in)                      // layer 1: direct convolution on the frames
                    rc = launch_conv_first(in, params + op.w, ones, zeros, c, B, h, w, op.cin, op.cout, 3, 1, 1, h, w, ACT_NONE, st);
                if (rc == ORBIT_ERR_UNSUPPORTED) {
                    rc = launch_im2col(in, buf[BUF_COL], B, h, w, op.cin, 3, 1, 1, 1, h, w, op.kpad, op.nchw_in, st);
                    if (rc) return rc;
                    rc = launch_pointwise_tcgen05(buf[BUF_COL], derived + op.w_split, ones, zeros, nullptr, nullptr, c, M, op.cout, op.kpad,
                                                  h * w, ACT_NONE, 3, st);
                    ++launches;
                }
                ++launches;
                break;
            }
            case OP_MAXPOOL: {
                const Op& conv = e->ops[i - 1];
                rc = launch_bn_relu_pool2_forward(saved + t.c_off[i - 1] * B, derived + conv.fold, derived + conv.fold + conv.cout,
                                                  saved + t.a_off[i] * B, B, h, w, op.cin, st);
                ++launches;
                break;
            }
            default: return ORBIT_ERR_UNSUPPORTED;
        }
        if (rc) return rc;
    }
    e->last_launches.store(launches);
    return ORBIT_OK;
}

extern "C" int orbit_engine_film_grad(const orbit_engine* e, const float* grad_params, float* grad_film, void* stream) {
    if (!e || !grad_params || !grad_film) return ORBIT_ERR_ARG;
    for (const FoldEntry& f : e->folds) {
        if (f.film_gamma >= 0)
            ORBIT_CUDA(cudaMemcpyAsync(grad_film + f.film_gamma, grad_params + f.gamma, sizeof(float) * f.channels, cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream));
        if (f.film_beta >= 0)
            ORBIT_CUDA(cudaMemcpyAsync(grad_film + f.film_beta, grad_params + f.beta, sizeof(float) * f.channels, cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream));
    }
    return e->lns.empty() ? ORBIT_OK : ORBIT_ERR_UNSUPPORTED;
}

// Set-encoder backward (model/set_encoders.py:81-120; every parameter is trainable: conv weight / bias, BatchNorm weight / bias;
// BatchNorm in eval mode, few_shot_recognisers.py:176-183). Gradients are ACCUMULATED into grad_params (layout of `params`).
static int set_encoder_backward(const orbit_engine* e, const TrainGeom& t, const float* params, const float* derived, float* tderived,
                                const float* saved, const float* frames, const float* dfeats, int B, float* grad_params, float** buf,
                                int64_t col_capacity, cudaStream_t st) {
    const float* ones = derived + e->ident;
    const float* zeros = derived + e->ident + e->max_c;
    int64_t launches = 0;
    int rc = ORBIT_OK;
    const float* dp = dfeats;          // gradient of the current pooled tensor (mode 1 for the first step: dfeats / HW)
    int mode = 1;
    for (int i = (int)e->ops.size() - 1; i >= 0; --i) {
        const Op& op = e->ops[i];
        if (op.kind != OP_CONV3) continue;
        const int h = t.in_h[i], w = t.in_w[i];
        const FoldEntry& f = e->folds[op.fold_idx];
        const float* c = saved + t.c_off[i] * B;
        float* dc = buf[BUF_E];
        // maxpool + ReLU + BatchNorm backward; partial sums at the head of the (idle) im2col buffer
        rc = launch_pool_bn_relu_backward(c, dp, derived + op.fold, derived + op.fold + op.cout, params + op.b, params + f.mean,
                                          params + f.var, f.eps, dc, buf[BUF_COL], grad_params + f.gamma, grad_params + f.beta,
                                          grad_params + op.b, B, h, w, op.cout, mode, st);
        if (rc) return rc;
        launches += 2;
        if (op.nchw_in) {
            // first conv (3 input channels): explicit im2col of the frames, weight gradient from the col matrix
            if (!frames) return ORBIT_ERR_ARG;
            rc = launch_im2col(frames, buf[BUF_COL], B, h, w, op.cin, 3, 1, 1, 1, h, w, op.kpad, 1, st);
            if (rc) return rc;
            // partial sums behind the col matrix: at most 56 M floats against the 116 M the buffer has left (M = B h w)
            rc = launch_col_wgrad(dc, buf[BUF_COL], buf[BUF_COL] + (int64_t)B * h * w * op.kpad, grad_params + op.w, (int64_t)B * h * w,
                                  op.cout, 9 * op.cin, op.kpad, st);
            if (rc) return rc;
            launches += 3;
        } else {
            const float* in = saved + t.a_off[i - 1] * B;       // pooled output of the previous block
            rc = launch_conv3_wgrad(dc, in, buf[BUF_COL], col_capacity, grad_params + op.w, B, h, w, op.cin, op.cout, st);
            if (rc) return rc;
            // data gradient: im2col(dc) x flipped / transposed weights -> gradient of the previous pooled tensor
            float* wd = tderived + t.wt_off[i];
            rc = launch_conv3_dgrad_weight(params + op.w, wd, op.cin, op.cout, st);
            if (rc) return rc;
            rc = launch_weight_split(wd, op.cin, 9 * op.cout, wd + (int64_t)op.cin * 9 * op.cout, st);
            if (rc) return rc;
            float* dprev = buf[e->ops[i - 1].out];               // the previous maxpool's output buffer
            rc = ORBIT_ERR_UNSUPPORTED;
            if (e->implicit_conv && op.cout % 64 == 0)
                rc = launch_conv3x3_tcgen05(dc, wd + (int64_t)op.cin * 9 * op.cout, ones, zeros, nullptr, dprev, B, h, w, op.cout, op.cin,
                                            ACT_NONE, st);
            if (rc == ORBIT_ERR_UNSUPPORTED) {
                rc = launch_im2col(dc, buf[BUF_COL], B, h, w, op.cout, 3, 1, 1, 1, h, w, 9 * op.cout, 0, st);
                if (rc) return rc;
                rc = launch_pointwise_tcgen05(buf[BUF_COL], wd + (int64_t)op.cin * 9 * op.cout, ones, zeros, nullptr, nullptr, dprev, B * h * w,
                                              op.cin, 9 * op.cout, h * w, ACT_NONE, 3, st);
            }
            if (rc) return rc;
            launches += 6;
            dp = dprev;
            mode = 0;
        }
    }
    e->last_launches.store(launches);
    return ORBIT_OK;
}

// grad_params: a blob with the layout of `params`; the gradients of the FiLM-site BatchNorm weight / bias are ACCUMULATED
// at the offsets of those parameters (everything else is left untouched). dfeats [num_frames, feat_dim].
extern "C" int orbit_engine_backward_train(const orbit_engine* e, const float* params, const float* derived, float* tderived,
                                           const float* saved, const float* frames, const float* dfeats, int num_frames, int height,
                                           int width, float* grad_params, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!e || !params || !derived || !tderived || !saved || !dfeats || !grad_params || !workspace) return ORBIT_ERR_ARG;
    if (num_frames <= 0 || height <= 0 || width <= 0) return ORBIT_ERR_ARG;
    const TrainGeom t = train_geometry(e, height, width);
    if (!t.ok) return ORBIT_ERR_UNSUPPORTED;
    const int B = num_frames;
    if (B > e->chunk_frames) return ORBIT_ERR_UNSUPPORTED;
    BufSizes bs;
    int rc = plan_buffers(e, height, width, &bs);
    if (rc) return rc;
    if (workspace_bytes < orbit_engine_workspace_bytes(e, height, width)) return ORBIT_ERR_WORKSPACE;
    float* buf[BUF_COUNT];
    workspace_buffers(e, bs, workspace, buf);
    cudaStream_t st = (cudaStream_t)stream;
    if (e->arch == ORBIT_ARCH_SET_ENCODER)
        return set_encoder_backward(e, t, params, derived, tderived, saved, frames, dfeats, B, grad_params, buf,
                                    bs.per_frame[BUF_COL] * e->chunk_frames, st);
    const float* ones = derived + e->ident;
    const float* zeros = derived + e->ident + e->max_c;
    int64_t launches = 0;
    bool from_pool = false;              // the gradient of the current tensor is dfeats / HW (global average pool)
    const float* skip = nullptr;         // gradient flowing around the block through the residual connection
    float* scratch = buf[BUF_E];         // per-channel partial sums of the FiLM-site reductions (free whenever a site is processed)
    auto film_grads = [&](const Op& op, float** gg, float** gb) {
        const FoldEntry& f = e->folds[op.fold_idx];
        const bool site = f.film_gamma >= 0;
        *gg = site ? grad_params + f.gamma : nullptr;
        *gb = site ? grad_params + f.beta : nullptr;
        return site;
    };
    for (int i = (int)e->ops.size() - 1; i >= 0; --i) {
        const Op& op = e->ops[i];
        const int h = t.in_h[i], w = t.in_w[i], ho = t.out_h[i], wo = t.out_w[i];
        const float* scale = op.fold >= 0 ? derived + op.fold : nullptr;
        const float* shift = op.fold >= 0 ? derived + op.fold + op.cout : nullptr;
        const FoldEntry* f = op.fold_idx >= 0 ? &e->folds[op.fold_idx] : nullptr;
        const float* c = t.c_off[i] >= 0 ? saved + t.c_off[i] * B : nullptr;
        switch (op.kind) {
            case OP_SPATIAL_MEAN: from_pool = true; break;
            case OP_PW: {
                const int M = B * h * w;
                const float* wt_split = tderived + t.wt_off[i] + (int64_t)op.cin * op.cout;
                if (op.act == ACT_SILU) {          // expand / conv_head: bn + SiLU backward, then dx = dc W
                    float *gg, *gb;
                    film_grads(op, &gg, &gb);
                    float* dc = buf[op.out];       // in place for the expand (its output gradient lives there), fresh for conv_head
                    rc = launch_bn_act_backward(c, from_pool ? dfeats : buf[op.out], nullptr, nullptr, scale, shift, params + f->mean,
                                                params + f->var, f->eps, dc, scratch == dc ? buf[BUF_D] : scratch, gg, gb, M, op.cout,
                                                ACT_SILU, from_pool ? 1 : 0, h * w, st);
                    if (rc) return rc;
                    from_pool = false;
                    rc = launch_pointwise_tcgen05(dc, wt_split, ones, zeros, nullptr, skip, buf[op.in], M, op.cin, op.cout, h * w, ACT_NONE, 3, st);
                    skip = nullptr;
                    launches += 2;
                } else {                           // gated project (no activation): dga = (dy scale3) W
                    const int other = op.out == BUF_X0 ? BUF_X1 : BUF_X0;          // the block-input buffer: free until dx is written
                    rc = launch_bn_act_backward(buf[op.out], buf[op.out], nullptr, nullptr, scale, shift, nullptr, nullptr, 0.f, buf[other],
                                                nullptr, nullptr, nullptr, M, op.cout, ACT_NONE, 0, h * w, st);
                    if (rc) return rc;
                    rc = launch_pointwise_tcgen05(buf[other], wt_split, ones, zeros, nullptr, nullptr, buf[op.in], M, op.cin, op.cout, h * w,
                                                  ACT_NONE, 3, st);
                    skip = op.res != BUF_NONE ? buf[op.out] : nullptr;
                    launches += 2;
                }
                break;
            }
            case OP_SE: {
                // dga lives in BUF_D (the project's input gradient); the depthwise op before holds a and the squeeze means
                rc = launch_se_backward(buf[BUF_D], saved + t.a_off[i - 1] * B, saved + t.m_off[i - 1] * B, params + op.w, params + op.b,
                                        params + op.w2, params + op.b2, buf[BUF_GATE], buf[BUF_PARTIAL], B, h * w, op.cin, op.se_reduce, st);
                launches += 2;
                break;
            }
            case OP_DW: {
                int pt, pl, d;
                same_geometry(h, op.k, op.stride, &d, &pt);
                same_geometry(w, op.k, op.stride, &d, &pl);
                float *gg, *gb;
                film_grads(op, &gg, &gb);
                // da = dga gate + dmean / HW (the SE op right after this one), then bn2 + SiLU backward, in place in BUF_D
                const float* gate = saved + t.g_off[i + 1] * B;
                rc = launch_bn_act_backward(c, buf[BUF_D], gate, buf[BUF_PARTIAL], scale, shift, params + f->mean, params + f->var, f->eps,
                                            buf[BUF_D], op.in == BUF_E ? buf[BUF_E] : buf[BUF_H], gg, gb, (int64_t)B * ho * wo, op.cout,
                                            ACT_SILU, 2, ho * wo, st);
                if (rc) return rc;
                rc = launch_dw_dgrad(buf[BUF_D], derived + op.dw_wt, buf[op.in], B, h, w, op.cin, ho, wo, op.k, op.stride, pt, pl, st);
                launches += 3;
                break;
            }
            case OP_STEM: {
                float *gg, *gb;
                if (film_grads(op, &gg, &gb))
                    rc = launch_bn_act_backward(c, buf[op.out], nullptr, nullptr, scale, shift, params + f->mean, params + f->var, f->eps,
                                                nullptr, scratch, gg, gb, (int64_t)B * ho * wo, op.cout, ACT_SILU, 0, ho * wo, st);
                launches += 2;
                break;
            }
            default: return ORBIT_ERR_UNSUPPORTED;
        }
        if (rc) return rc;
    }
    e->last_launches.store(launches);
    return ORBIT_OK;
}
