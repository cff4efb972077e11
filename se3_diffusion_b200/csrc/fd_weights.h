// Parameter schema (the reference's 282-entry state_dict) and the packed device layout.
#pragma once
#include <string>
#include <vector>
#include <stdint.h>
#include "fd_common.cuh"

namespace fd {

struct ParamDesc {
  std::string name;
  int ndim;
  int64_t dim[2];
  int64_t numel() const { return ndim == 1 ? dim[0] : dim[0] * dim[1]; }
};

inline const std::vector<ParamDesc>& param_schema() {
  static std::vector<ParamDesc> s;
  if (!s.empty()) return s;
  auto lin = [&](const std::string& n, int64_t o, int64_t i) {
    s.push_back({n + ".weight", 2, {o, i}});
    s.push_back({n + ".bias", 1, {o, 0}});
  };
  auto ln = [&](const std::string& n, int64_t c) {
    s.push_back({n + ".weight", 1, {c, 0}});
    s.push_back({n + ".bias", 1, {c, 0}});
  };
  const std::string e = "embedding_layer.";
  lin(e + "node_embedder.0", 256, 65); lin(e + "node_embedder.2", 256, 256); lin(e + "node_embedder.4", 256, 256);
  ln(e + "node_embedder.5", 256);
  lin(e + "edge_embedder.0", 128, 120); lin(e + "edge_embedder.2", 128, 128); lin(e + "edge_embedder.4", 128, 128);
  ln(e + "edge_embedder.5", 128);
  const std::string t = "score_model.trunk.";
  for (int b = 0; b < NBLK; ++b) {
    const std::string sb = std::to_string(b);
    s.push_back({t + "ipa_" + sb + ".head_weights", 1, {H, 0}});
    lin(t + "ipa_" + sb + ".linear_q", PROJ_Q, C_S);
    lin(t + "ipa_" + sb + ".linear_kv", PROJ_KV, C_S);
    lin(t + "ipa_" + sb + ".linear_q_points", PROJ_QP, C_S);
    lin(t + "ipa_" + sb + ".linear_kv_points", PROJ_KVP, C_S);
    lin(t + "ipa_" + sb + ".linear_b", H, C_Z);
    lin(t + "ipa_" + sb + ".down_z", C_Z / 4, C_Z);
    lin(t + "ipa_" + sb + ".linear_out", C_S, IPA_FEAT);
    lin(t + "ipa_" + sb + ".linear_rbf", 1, 20);
    ln(t + "ipa_ln_" + sb, C_S);
    lin(t + "skip_embed_" + sb, C_SKIP, C_S);
    for (int l = 0; l < TF_LAYERS; ++l) {
      const std::string p = t + "seq_tfmr_" + sb + ".layers." + std::to_string(l) + ".";
      s.push_back({p + "self_attn.in_proj_weight", 2, {3 * TF_D, TF_D}});
      s.push_back({p + "self_attn.in_proj_bias", 1, {3 * TF_D, 0}});
      lin(p + "self_attn.out_proj", TF_D, TF_D);
      lin(p + "linear1", TF_D, TF_D); lin(p + "linear2", TF_D, TF_D);
      ln(p + "norm1", TF_D); ln(p + "norm2", TF_D);
    }
    lin(t + "post_tfmr_" + sb, C_S, TF_D);
    for (int k = 1; k <= 3; ++k) lin(t + "node_transition_" + sb + ".linear_" + std::to_string(k), C_S, C_S);
    ln(t + "node_transition_" + sb + ".ln", C_S);
    lin(t + "bb_update_" + sb + ".linear", 6, C_S);
    if (b < NBLK - 1) {
      const std::string p = t + "edge_transition_" + sb + ".";
      lin(p + "initial_embed", C_Z, C_S);
      lin(p + "trunk.0", ET_HID, ET_HID); lin(p + "trunk.2", ET_HID, ET_HID);
      lin(p + "final_layer", C_Z, ET_HID);
      ln(p + "layer_norm", C_Z);
    }
  }
  const std::string p = "score_model.torsion_pred.";
  lin(p + "linear_1", C_S, C_S); lin(p + "linear_2", C_S, C_S); lin(p + "linear_3", C_S, C_S);
  lin(p + "linear_final", 2, C_S);
  return s;
}

// Device-resident packed weights (all fp32; the tensor-core path keeps additional bf16 hi/lo images, see fd_tc.cuh).
struct Lin { const float* w = nullptr; const float* b = nullptr; };
struct LNp { const float* g = nullptr; const float* b = nullptr; };

struct TfLayer { Lin in_proj, out_proj, lin1, lin2; LNp norm1, norm2; };

struct BlockW {
  Lin proj;                  // [6816][256] = q | kv | q_points | kv_points
  const float* Wb = nullptr; const float* bb = nullptr;   // linear_b [8][128], [8]
  const float* gamma = nullptr;                           // softplus(head_weights)·sqrt(1/(3·PQ·9/2)) [8]
  const float* WdT = nullptr; const float* bd = nullptr;  // down_z transposed [128][32], [32]
  Lin down_bd;                                            // down_z as one block-diagonal linear over all heads: [256][1024], bias [256] (tensor-core path)
  Lin out;                   // linear_out [256][2688]
  LNp ipa_ln;
  Lin skip;                  // [64][256]
  TfLayer tf[TF_LAYERS];
  Lin post;                  // [256][320]
  Lin tr1, tr2, tr3; LNp tr_ln;
  Lin bbu;                   // [6][256]
  // edge transition (blocks 0..2)
  Lin et_init;               // [128][256]
  Lin et_node;               // [1024][128]: rows P(384: W1[:,128:256], +b1) | Q(384: W1[:,256:384]) | U(128: Wf[:,128:256], +bf) | V(128: Wf[:,256:384])
  const float* et_w1z = nullptr;   // [384][128] = W1[:, 0:128]
  Lin et_w2;                 // [384][384] + b2
  const float* et_wfh = nullptr;   // [128][384] = Wf (applied to h2)
  const float* et_wfz = nullptr;   // [128][128] = Wf[:, 0:128]
  LNp et_ln;
};

struct Weights {
  Lin ne0, ne2, ne4; LNp ne_ln;          // node embedder (ne0 K padded 65 -> 68)
  const float* ee_w0a = nullptr;         // [33][128] transposed W0[:, 0:33]
  const float* ee_w0c = nullptr;         // [33][128] transposed W0[:, 33:66]
  const float* ee_w0r = nullptr;         // [32][128] transposed W0[:, 66:98]
  const float* ee_D = nullptr;           // [23][128] transposed W0[:, 98:120] + zero row
  const float* ee_b0 = nullptr;
  float* ee_T = nullptr;                 // [2*REL_DMAX+1][128] rel-offset table (built on device)
  Lin ee2, ee4; LNp ee_ln;
  BlockW blk[NBLK];
  Lin tor1, tor2, torf;
};

}  // namespace fd
