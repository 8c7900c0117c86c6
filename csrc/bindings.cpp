// TORCH_LIBRARY registration of every vescale_b200 native op (loaded with torch.ops.load_library).
#include <ATen/ATen.h>
#include <torch/library.h>

#include <tuple>

// elementwise.cu
std::tuple<at::Tensor, at::Tensor> rms_norm_fwd(const at::Tensor& x, const at::Tensor& w, double eps);
std::tuple<at::Tensor, at::Tensor, at::Tensor> add_rms_norm_fwd(const at::Tensor& a, const at::Tensor& b, const at::Tensor& w, double eps);
std::tuple<at::Tensor, at::Tensor> rms_norm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w, const at::Tensor& rstd);
std::tuple<at::Tensor, at::Tensor> add_rms_norm_bwd(const at::Tensor& dy, const at::Tensor& dh, const at::Tensor& h, const at::Tensor& w, const at::Tensor& rstd);
at::Tensor swiglu_fwd(const at::Tensor& gu);
at::Tensor swiglu_bwd(const at::Tensor& dy, const at::Tensor& gu);
void rope_qk_(at::Tensor qkv, const at::Tensor& cos, const at::Tensor& sin, int64_t S, int64_t n_q, int64_t n_kv, int64_t D, double sign);
at::Tensor cross_entropy_fwd_bwd_(at::Tensor logits, const at::Tensor& target, const at::Tensor& n_valid, int64_t ignore_index);
void sumsq_accumulate(const at::Tensor& g, at::Tensor out, double scale);
void fused_adamw_(at::Tensor master, at::Tensor m, at::Tensor v, const at::Tensor& g, at::Tensor p_out, const at::Tensor& wd_table,
                  const at::Tensor& coef, double lr, double b1, double b2, double eps, double wd, double bc1, double bc2, double gscale);
// gemm_sm100.cu
void mx_quantize(const at::Tensor& x, at::Tensor q, at::Tensor sf);
void mxfp8_gemm_nt(const at::Tensor& a_q, const at::Tensor& sfa, const at::Tensor& b_q, const at::Tensor& sfb, at::Tensor c);
void gemm_nt(const at::Tensor& a, const at::Tensor& b, at::Tensor c, bool accumulate, int64_t variant);
void gemm_nn(const at::Tensor& a, const at::Tensor& b, at::Tensor c);
void gemm_tn(const at::Tensor& a, const at::Tensor& b, at::Tensor c, bool accumulate);

// symm_comm.cu
void attn_fwd(const at::Tensor& qkv, at::Tensor out, at::Tensor lse, int64_t n_q, int64_t n_kv, double softmax_scale, int64_t variant);
void attn_bwd(const at::Tensor& qkv, const at::Tensor& out, const at::Tensor& dout, const at::Tensor& lse, at::Tensor dqkv, at::Tensor dvec, at::Tensor dq_acc,
              int64_t n_q, int64_t n_kv, double softmax_scale);
void ragged_norm_partial(const at::Tensor& x, at::Tensor out, int64_t rows, int64_t C, int64_t mode, int64_t p);
void gemm_set_sched(int64_t mode);
void attn_set_bwd_variant(int64_t v);
int64_t attn_get_bwd_variant();
int64_t gemm_get_sched();
void symm_signal(std::vector<int64_t> pad_ptrs, int64_t rank, int64_t slot, int64_t epoch);
void symm_wait(int64_t my_pad, int64_t world, int64_t slot, int64_t epoch);
void symm_all_gather_ce(std::vector<int64_t> shard_ptrs, at::Tensor full, int64_t shard_bytes, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot,
                        int64_t epoch, int64_t lo_bytes, int64_t hi_bytes);
void symm_all_gather(std::vector<int64_t> shard_ptrs, at::Tensor full, int64_t shard_bytes, int64_t rank, std::vector<int64_t> pad_ptrs,
                     int64_t slot, int64_t epoch, int64_t num_ctas, int64_t range_mode, int64_t range_lo_bytes, int64_t range_hi_bytes);
void symm_reduce_scatter(std::vector<int64_t> grad_ptrs, at::Tensor out, c10::optional<at::Tensor> sumsq, int64_t shard_elems, int64_t rank,
                         double scale, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, int64_t multicast_ptr, int64_t num_ctas);
void symm_rs_adamw(std::vector<int64_t> grad_ptrs, at::Tensor master, at::Tensor m, at::Tensor v, at::Tensor p_out, const at::Tensor& wd_table,
                   const at::Tensor& coef, c10::optional<at::Tensor> sumsq, int64_t rank, double scale, std::vector<int64_t> pad_ptrs, int64_t slot,
                   int64_t epoch, double lr, double b1, double b2, double eps, double wd, double bc1, double bc2, int64_t num_ctas);

// symm_collectives.cu
void symm_reduce_scatter_t(std::vector<int64_t> buf_ptrs, int64_t multicast_ptr, at::Tensor out, int64_t slice_numel, int64_t row_numel, int64_t out_row_stride,
                           int64_t dtype_code, double scale, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter,
                           int64_t num_ctas);
void symm_all_reduce(std::vector<int64_t> buf_ptrs, int64_t multicast_ptr, c10::optional<at::Tensor> out, int64_t numel, int64_t dtype_code, double scale,
                     int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter, int64_t num_ctas);
void symm_a2a_permute(const at::Tensor& src, std::vector<int64_t> dst_ptrs, std::vector<int64_t> n, std::vector<int64_t> ss, std::vector<int64_t> ds,
                      int64_t src_peer_stride, int64_t dst_rank_stride, int64_t vec_bytes, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot,
                      int64_t epoch, at::Tensor counter, int64_t num_ctas);
void symm_put_segments(const at::Tensor& src, std::vector<int64_t> dst_ptrs, const at::Tensor& table, int64_t vec_bytes, int64_t total_vecs, int64_t rank,
                       std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter, int64_t num_ctas);
at::Tensor symm_vocab_ce_(at::Tensor logits, const at::Tensor& target, const at::Tensor& n_valid, int64_t vocab_start, int64_t ignore_index,
                          std::vector<int64_t> stats_ptrs, int64_t rank, int64_t epoch, int64_t max_rows, int64_t max_ctas);

// gemm_fused_tp.cu
void ag_gemm(const at::Tensor& x_local, std::vector<int64_t> x_ptrs, const at::Tensor& w, at::Tensor x_full, at::Tensor y, at::Tensor arrive,
             std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch);
void gemm_rs(const at::Tensor& x, const at::Tensor& w, at::Tensor y, std::vector<int64_t> staging_ptrs, at::Tensor done, std::vector<int64_t> flag_ptrs,
             int64_t rank, int64_t epoch);

void wag_gemm(const at::Tensor& x, at::Tensor w_full, std::vector<int64_t> shard_ptrs, std::vector<int64_t> row_bounds, at::Tensor y, at::Tensor arrive,
              std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch, bool wait_peers);
void grouped_gemm_nt(const at::Tensor& a, const at::Tensor& b, at::Tensor c, const at::Tensor& tile_expert, int64_t expert_n);
// moe_dispatch.cu
void moe_exchange_counts(const at::Tensor& my_counts, std::vector<int64_t> counts_all_ptrs, std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch);
void moe_plan(const at::Tensor& counts_all, at::Tensor recv_seg_start, at::Tensor send_off, at::Tensor tile_expert, at::Tensor expert_rows,
              at::Tensor total_rows, int64_t rank, int64_t tile_rows);
void moe_dispatch_put(const at::Tensor& rows, const at::Tensor& row_expert, const at::Tensor& row_pos, const at::Tensor& send_off,
                      std::vector<int64_t> recv_ptrs, std::vector<int64_t> flag_ptrs, at::Tensor done_counter, int64_t E, int64_t rank, int64_t epoch);
void moe_wait(int64_t my_flags, int64_t W, int64_t base, int64_t epoch);
void moe_signal(std::vector<int64_t> flag_ptrs, int64_t base, int64_t rank, int64_t epoch);
void moe_combine_get(at::Tensor out, c10::optional<at::Tensor> gate, const at::Tensor& slot_rank, const at::Tensor& slot_row, std::vector<int64_t> src_ptrs,
                     int64_t k);

// philox_shard.cu
void philox_fill_box(at::Tensor local, std::vector<int64_t> size, std::vector<int64_t> goff, std::vector<int64_t> gstride, std::vector<int64_t> lstride,
                     int64_t lbase, int64_t seed, int64_t offset, bool normal, double a, double b);

void philox_dropout_box(const at::Tensor& x, at::Tensor out, at::Tensor mask, std::vector<int64_t> size, std::vector<int64_t> goff, std::vector<int64_t> gstride,
                        std::vector<int64_t> lstride, int64_t lbase, int64_t seed, int64_t offset, double p);

TORCH_LIBRARY(vescale_b200, m) {
  m.def("philox_dropout_box(Tensor x, Tensor(a!) out, Tensor(b!) mask, int[] size, int[] goff, int[] gstride, int[] lstride, int lbase, int seed, int offset, float p) -> ()");
  m.def("rms_norm_fwd(Tensor x, Tensor w, float eps) -> (Tensor, Tensor)");
  m.def("add_rms_norm_fwd(Tensor a, Tensor b, Tensor w, float eps) -> (Tensor, Tensor, Tensor)");
  m.def("rms_norm_bwd(Tensor dy, Tensor x, Tensor w, Tensor rstd) -> (Tensor, Tensor)");
  m.def("add_rms_norm_bwd(Tensor dy, Tensor dh, Tensor h, Tensor w, Tensor rstd) -> (Tensor, Tensor)");
  m.def("swiglu_fwd(Tensor gu) -> Tensor");
  m.def("swiglu_bwd(Tensor dy, Tensor gu) -> Tensor");
  m.def("rope_qk_(Tensor(a!) qkv, Tensor cos, Tensor sin, int S, int n_q, int n_kv, int D, float sign) -> ()");
  m.def("cross_entropy_fwd_bwd_(Tensor(a!) logits, Tensor target, Tensor n_valid, int ignore_index) -> Tensor");
  m.def("sumsq_accumulate(Tensor g, Tensor(a!) out, float scale) -> ()");
  m.def("fused_adamw_(Tensor(a!) master, Tensor(b!) m, Tensor(c!) v, Tensor g, Tensor(d!) p_out, Tensor wd_table, Tensor coef, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale) -> ()");
  m.def("mx_quantize(Tensor x, Tensor(a!) q, Tensor(b!) sf) -> ()");
  m.def("mxfp8_gemm_nt(Tensor a_q, Tensor sfa, Tensor b_q, Tensor sfb, Tensor(a!) c) -> ()");
  m.def("gemm_nt(Tensor a, Tensor b, Tensor(a!) c, bool accumulate, int variant=0) -> ()");
  m.def("gemm_nn(Tensor a, Tensor b, Tensor(a!) c) -> ()");
  m.def("philox_fill_box(Tensor(a!) local, int[] size, int[] goff, int[] gstride, int[] lstride, int lbase, int seed, int offset, bool normal, float a, float b) -> ()");
  m.def("grouped_gemm_nt(Tensor a, Tensor b, Tensor(a!) c, Tensor tile_expert, int expert_n) -> ()");
  m.def("moe_exchange_counts(Tensor my_counts, int[] counts_all_ptrs, int[] flag_ptrs, int rank, int epoch) -> ()");
  m.def("moe_plan(Tensor counts_all, Tensor(a!) recv_seg_start, Tensor(b!) send_off, Tensor(c!) tile_expert, Tensor(d!) expert_rows, Tensor(e!) total_rows, int rank, int tile_rows) -> ()");
  m.def("moe_dispatch_put(Tensor rows, Tensor row_expert, Tensor row_pos, Tensor send_off, int[] recv_ptrs, int[] flag_ptrs, Tensor(a!) done_counter, int E, int rank, int epoch) -> ()");
  m.def("moe_wait(int my_flags, int W, int base, int epoch) -> ()", &moe_wait);
  m.def("moe_signal(int[] flag_ptrs, int base, int rank, int epoch) -> ()", &moe_signal);
  m.def("moe_combine_get(Tensor(a!) out, Tensor? gate, Tensor slot_rank, Tensor slot_row, int[] src_ptrs, int k) -> ()");
  m.def("gemm_tn(Tensor a, Tensor b, Tensor(a!) c, bool accumulate) -> ()");
  m.def("ag_gemm(Tensor x_local, int[] x_ptrs, Tensor w, Tensor(a!) x_full, Tensor(b!) y, Tensor(c!) arrive, int[] flag_ptrs, int rank, int epoch) -> ()");
  m.def("gemm_rs(Tensor x, Tensor w, Tensor(a!) y, int[] staging_ptrs, Tensor(b!) done, int[] flag_ptrs, int rank, int epoch) -> ()");
  m.def("attn_fwd(Tensor qkv, Tensor(a!) out, Tensor(b!) lse, int n_q, int n_kv, float softmax_scale, int variant=0) -> ()");
  m.def("attn_bwd(Tensor qkv, Tensor out, Tensor dout, Tensor lse, Tensor(a!) dqkv, Tensor(b!) dvec, Tensor(c!) dq_acc, int n_q, int n_kv, float softmax_scale) -> ()");
  m.def("ragged_norm_partial(Tensor x, Tensor(a!) out, int rows, int C, int mode, int p) -> ()");
  m.def("gemm_set_sched(int mode) -> ()", &gemm_set_sched);
  m.def("attn_set_bwd_variant(int v) -> ()", &attn_set_bwd_variant);
  m.def("attn_get_bwd_variant() -> int", &attn_get_bwd_variant);
  m.def("gemm_get_sched() -> int", &gemm_get_sched);
  m.def("symm_signal(int[] pad_ptrs, int rank, int slot, int epoch) -> ()", &symm_signal);
  m.def("symm_wait(int my_pad, int world, int slot, int epoch) -> ()", &symm_wait);
  m.def("symm_all_gather_ce(int[] shard_ptrs, Tensor(a!) full, int shard_bytes, int rank, int[] pad_ptrs, int slot, int epoch, int lo_bytes=0, int hi_bytes=0) -> ()");
  m.def("symm_all_gather(int[] shard_ptrs, Tensor(a!) full, int shard_bytes, int rank, int[] pad_ptrs, int slot, int epoch, int num_ctas, int range_mode=0, int range_lo_bytes=0, int range_hi_bytes=0) -> ()");
  m.def("symm_reduce_scatter(int[] grad_ptrs, Tensor(a!) out, Tensor(b!)? sumsq, int shard_elems, int rank, float scale, int[] pad_ptrs, int slot, int epoch, int multicast_ptr, int num_ctas) -> ()");
  m.def("symm_rs_adamw(int[] grad_ptrs, Tensor(a!) master, Tensor(b!) m, Tensor(c!) v, Tensor(d!) p_out, Tensor wd_table, Tensor coef, Tensor(e!)? sumsq, int rank, float scale, int[] pad_ptrs, int slot, int epoch, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, int num_ctas) -> ()");
  m.def("wag_gemm(Tensor x, Tensor(a!) w_full, int[] shard_ptrs, int[] row_bounds, Tensor(b!) y, Tensor(c!) arrive, int[] flag_ptrs, int rank, int epoch, bool wait_peers=True) -> ()");
  m.def("symm_all_reduce(int[] buf_ptrs, int multicast_ptr, Tensor(a!)? out, int numel, int dtype_code, float scale, int rank, int[] pad_ptrs, int slot, int epoch, Tensor(b!) counter, int num_ctas) -> ()");
  m.def("symm_reduce_scatter_t(int[] buf_ptrs, int multicast_ptr, Tensor(a!) out, int slice_numel, int row_numel, int out_row_stride, int dtype_code, float scale, int rank, int[] pad_ptrs, int slot, int epoch, Tensor(b!) counter, int num_ctas) -> ()");
  m.def("symm_a2a_permute(Tensor src, int[] dst_ptrs, int[] n, int[] ss, int[] ds, int src_peer_stride, int dst_rank_stride, int vec_bytes, int rank, int[] pad_ptrs, int slot, int epoch, Tensor(a!) counter, int num_ctas) -> ()");
  m.def("symm_put_segments(Tensor src, int[] dst_ptrs, Tensor table, int vec_bytes, int total_vecs, int rank, int[] pad_ptrs, int slot, int epoch, Tensor(a!) counter, int num_ctas) -> ()");
  m.def("symm_vocab_ce_(Tensor(a!) logits, Tensor target, Tensor n_valid, int vocab_start, int ignore_index, int[] stats_ptrs, int rank, int epoch, int max_rows, int max_ctas) -> Tensor");
}

TORCH_LIBRARY_IMPL(vescale_b200, CUDA, m) {
  m.impl("philox_dropout_box", &philox_dropout_box);
  m.impl("wag_gemm", &wag_gemm);
  m.impl("symm_all_reduce", &symm_all_reduce);
  m.impl("symm_reduce_scatter_t", &symm_reduce_scatter_t);
  m.impl("symm_a2a_permute", &symm_a2a_permute);
  m.impl("symm_put_segments", &symm_put_segments);
  m.impl("symm_vocab_ce_", &symm_vocab_ce_);
  m.impl("rms_norm_fwd", &rms_norm_fwd);
  m.impl("add_rms_norm_fwd", &add_rms_norm_fwd);
  m.impl("rms_norm_bwd", &rms_norm_bwd);
  m.impl("add_rms_norm_bwd", &add_rms_norm_bwd);
  m.impl("swiglu_fwd", &swiglu_fwd);
  m.impl("swiglu_bwd", &swiglu_bwd);
  m.impl("rope_qk_", &rope_qk_);
  m.impl("cross_entropy_fwd_bwd_", &cross_entropy_fwd_bwd_);
  m.impl("sumsq_accumulate", &sumsq_accumulate);
  m.impl("fused_adamw_", &fused_adamw_);
  m.impl("gemm_nt", &gemm_nt);
  m.impl("mxfp8_gemm_nt", &mxfp8_gemm_nt);
  m.impl("mx_quantize", &mx_quantize);
  m.impl("gemm_nn", &gemm_nn);
  m.impl("philox_fill_box", &philox_fill_box);
  m.impl("grouped_gemm_nt", &grouped_gemm_nt);
  m.impl("moe_exchange_counts", &moe_exchange_counts);
  m.impl("moe_plan", &moe_plan);
  m.impl("moe_dispatch_put", &moe_dispatch_put);
  m.impl("moe_combine_get", &moe_combine_get);
  m.impl("gemm_tn", &gemm_tn);
  m.impl("ag_gemm", &ag_gemm);
  m.impl("gemm_rs", &gemm_rs);
  m.impl("ragged_norm_partial", &ragged_norm_partial);
  m.impl("attn_fwd", &attn_fwd);
  m.impl("attn_bwd", &attn_bwd);
  m.impl("symm_all_gather", &symm_all_gather);
  m.impl("symm_all_gather_ce", &symm_all_gather_ce);
  m.impl("symm_reduce_scatter", &symm_reduce_scatter);
  m.impl("symm_rs_adamw", &symm_rs_adamw);
}
