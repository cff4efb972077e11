// tcgen05 tensor-core edge kernels (FD_PREC_BF16X3 / FD_PREC_BF16): interface used by fd_engine.cu.
// STAGE 0 (this file): interface only; the kernels land in the next milestone.  Selecting a tensor-core precision
// fails loudly (FD_ESTATE) instead of silently falling back to the fp32 path.
#pragma once
#include <map>
#include <string>
#include "fd_common.cuh"

namespace fd {

struct TcWeights { bool ready = false; };
struct TcWorkspace { char* base = nullptr; };

inline int tc_init(int sm_count) { (void)sm_count; return 0; }
inline void tc_free_weights(TcWeights&) {}
inline int tc_pack_weights(TcWeights&, const std::map<std::string, const float*>&, cudaStream_t) { return 0; }
inline size_t tc_workspace_bytes(int, int) { return 0; }
inline void tc_bind_workspace(TcWorkspace&, char*, int, int) {}
inline int tc_unavailable() { return -4; }
inline int tc_edge_embed(const TcWeights&, TcWorkspace&, int, const float*, const float*, const float*, const float*, const int*,
                         const float*, const float*, int, int, cudaStream_t, long long*) { return tc_unavailable(); }
inline void tc_export_z(TcWorkspace&, float*, int, int, cudaStream_t) {}
inline int tc_ipa_edge(TcWorkspace&, float*, const float*, const float*, const float*, const float*, const float*, const float*,
                       const float*, const float*, float*, int, int, int, int, cudaStream_t, long long*) { return tc_unavailable(); }
inline int tc_edge_transition(const TcWeights&, TcWorkspace&, int, int, const float*, const float*, const float*, const float*,
                              const float*, int, int, cudaStream_t, long long*) { return tc_unavailable(); }

}  // namespace fd
