// Host-only shape planner: mirrors, per stream, every length-dependent decision of the reference's
// streaming path so that one push of many ragged streams becomes a handful of flat descriptor arrays.
// Nothing here depends on data values, only on sample counts -- which is what lets the whole
// frontend/encoder schedule be computed on the host without a device round trip.
//
// Mirrors: speechcatcher/speech2text_streaming.py:300-400 (waveform buffering, trimming)
//          speechcatcher/model/encoder/contextual_block_transformer_encoder.py:278-419, 500-522
//          speechcatcher/beam_search/beam_search.py:551-634 (encoder skip, block trigger, final block)
#pragma once
#include <vector>
#include "kernels.h"

namespace scb {

struct StreamHost {
  // frontend state (speech2text_streaming.py: frontend_states["waveform_buffer"])
  bool fe_state = false;
  int wbuf_n = 0;
  // encoder state dict (contextual_block_transformer_encoder.py:410-417)
  bool enc_state = false;
  int feat_carry = 0;      // buffer_before_downsampling frames
  int sub_n = 0;           // buffer_after_downsampling frames
  int n_proc = 0;          // n_processed_blocks
  bool has_prev_addin = false;
  bool has_past_ctx = false;
  // search state (beam_search.py:316-327)
  int enc_len = 0;
  int processed_block = 0;
};

struct StreamPush {
  bool has_fd = false;
  FrontendDesc fd{};
  bool called = false;       // process_block is invoked for this stream
  int n_feat = 0;
  bool run_sub = false;
  SubDesc sd{};
  bool feat_carry_move = false;
  bool sub_carry_move = false;
  int sub_carry_src = 0, sub_carry_n = 0;
  std::vector<BlockDesc> blocks;   // prev_blk is relative to this vector
  int enc_t0 = 0;                  // first new encoder frame
  int n_enc_out = 0;
  std::vector<int> dq_T, dq_final;
  int error = 0;                   // non-zero: shape the reference itself cannot process
};

inline int conv_out_len(int t) { return (((t - 3) / 2 + 1) - 3) / 2 + 1; }

class Planner {
 public:
  explicit Planner(int n_streams) : hs_(n_streams) {}
  void reset(int s) { hs_[s] = StreamHost(); }
  const StreamHost& state(int s) const { return hs_[s]; }
  void restore(int s, const StreamHost& h) { hs_[s] = h; }      // roll a failed push back (engine.cu push_impl)

  StreamPush push(int s, int n_new, bool is_final) {
    StreamHost& h = hs_[s];
    StreamPush p;
    const int WIN = 400, HOP = 160;
    // ---------------- frontend (speech2text_streaming.py:300-400)
    const bool had_state = h.fe_state;
    const int n_prev = had_state ? h.wbuf_n : 0;
    const int total = n_prev + n_new;
    p.has_fd = true;
    p.fd.stream = s; p.fd.n_prev = n_prev; p.fd.n_new = n_new;
    if (total <= WIN && !is_final) {                 // :306-319 buffer and return
      p.fd.slab = 0; p.fd.n_frames = 0; p.fd.emit0 = p.fd.emit1 = 0; p.fd.new_buf = total;
      h.fe_state = true; h.wbuf_n = total;
      return p;
    }
    int slab, new_buf;
    if (is_final) { slab = total <= WIN ? WIN : total; new_buf = 0; }
    else {
      const int nfr = (total - (WIN - HOP)) / HOP, nres = (total - (WIN - HOP)) % HOP;
      slab = (WIN - HOP) + nfr * HOP;
      new_buf = (WIN - HOP) + nres;
    }
    const int F = 1 + slab / HOP;
    int e0 = 0, e1 = F;
    bool emit = true;
    if (is_final) { if (had_state && F > 2) e0 = 2; }
    else if (!had_state) { if (F > 2) e1 = F - 2; }
    else { if (F > 4) { e0 = 2; e1 = F - 2; } else emit = false; }     // :384-389
    p.fd.slab = slab; p.fd.n_frames = F; p.fd.new_buf = new_buf;
    h.fe_state = !is_final; h.wbuf_n = new_buf;
    if (!emit) { p.fd.emit0 = p.fd.emit1 = 0; return p; }
    p.fd.emit0 = e0; p.fd.emit1 = e1;
    p.called = true;
    p.n_feat = e1 - e0;
    p.fd.feat_off = h.enc_state ? h.feat_carry : 0;
    encoder_and_trigger(h, p, s, is_final);
    return p;
  }

  // Pre-computed features instead of a waveform (speech2text_streaming.py:438-450: the frontend is skipped and
  // process_block is always called): `n_feat` new feature frames go to featbuf[s] at row `feat_off`.
  StreamPush push_features(int s, int n_feat, bool is_final) {
    StreamHost& h = hs_[s];
    StreamPush p;
    p.has_fd = false;
    p.called = true;
    p.n_feat = n_feat;
    p.fd.stream = s;
    p.fd.feat_off = h.enc_state ? h.feat_carry : 0;
    encoder_and_trigger(h, p, s, is_final);
    return p;
  }

 private:
  // everything after the frontend: encoder buffering / blocks and the decode-block trigger
  void encoder_and_trigger(StreamHost& h, StreamPush& p, int s, bool is_final) {
    const int carry = p.fd.feat_off;
    // ---------------- encoder (contextual_block_transformer_encoder.py:278-419)
    if (p.n_feat >= 3) {                              // beam_search.py:551: skip the encoder otherwise
      const int T = carry + p.n_feat;
      int t_in = 0, t2 = 0;
      bool produce = true;
      if (!is_final) {
        const int n_samples = T / 4 - 1;
        if (n_samples < 2) { h.feat_carry = T; h.enc_state = true; produce = false; }
        else {
          const int n_res = T % 4 + 8;
          t_in = n_samples * 4; t2 = n_samples - 1;
          p.feat_carry_move = true;
          p.sd.carry_src = T - n_res; p.sd.carry_n = n_res;
          h.feat_carry = n_res; h.enc_state = true;
        }
      } else {
        if (T < 7) { p.error = 1; return; }        // conv2d would raise in the reference
        t_in = T; t2 = conv_out_len(T);
        h.feat_carry = 0;
      }
      if (produce) {
        p.run_sub = true;
        p.sd.stream = s; p.sd.t_in = t_in; p.sd.t1 = (t_in - 3) / 2 + 1; p.sd.t2 = t2; p.sd.sub_off = h.sub_n;
        const int totf = h.sub_n + t2;
        int bn = 0;
        bool blocks = true, short_path = false;
        if (is_final) {
          short_path = (h.n_proc == 0 && totf <= kBlock);
          if (!short_path) {
            bn = (totf - (kBlock - kHopB - kLook) - kLook + kHopB - 1) / kHopB;      // ceil((tot-24)/16)
            if (totf - 24 <= 0) bn = 0;
            if (bn <= 0) { p.error = 2; return; }   // the reference indexes an empty block tensor here
          }
          h.sub_n = 0;
        } else {
          if (totf <= kBlock) { h.sub_n = totf; blocks = false; }
          else {
            bn = (totf - (kBlock - kHopB)) / kHopB;
            const int res = totf - kHopB * bn;
            p.sub_carry_move = true; p.sub_carry_src = totf - res; p.sub_carry_n = res;
            h.sub_n = res;
          }
        }
        if (blocks) {
          p.enc_t0 = h.enc_len;
          if (short_path) {
            BlockDesc b{};
            b.stream = s; b.sub_start = 0; b.clen = totf; b.pe_frame_off = 0; b.pe_ctx_off = 0; b.prev_blk = -1;
            b.is_last = 1; b.out_slot0 = 0; b.out_count = totf; b.out_t0 = h.enc_len; b.n_rows = totf;
            b.short_path = 1;
            p.blocks.push_back(b);
            p.n_enc_out = totf;
          } else {
            const int off = kBlock - kLook - kHopB;   // 8
            const bool first = h.n_proc == 0;
            const int y_len = is_final ? (first ? totf : totf - off) : bn * kHopB + (first ? off : 0);
            for (int i = 0; i < bn; ++i) {
              BlockDesc b{};
              b.stream = s; b.sub_start = i * kHopB; b.clen = std::min(kBlock, totf - i * kHopB);
              b.pe_frame_off = i * kHopB + kHopB * h.n_proc; b.pe_ctx_off = i + h.n_proc;
              b.prev_blk = i - 1; b.has_prev_addin = h.has_prev_addin; b.has_past_ctx = h.has_past_ctx;
              b.is_last = (i == bn - 1); b.n_rows = kSlots; b.short_path = 0;
              const int cur = i * kHopB + (first ? off : 0);
              const int cl = (i == bn - 1 && is_final) ? std::min(kBlock - off, y_len - cur) : kHopB;
              if (first && i == 0) { b.out_slot0 = 1; b.out_count = off + cl; b.out_t0 = h.enc_len; }
              else { b.out_slot0 = 1 + off; b.out_count = cl; b.out_t0 = h.enc_len + cur; }
              p.blocks.push_back(b);
            }
            p.n_enc_out = y_len;
            h.n_proc += bn; h.has_prev_addin = true; h.has_past_ctx = true;
          }
          h.enc_len += p.n_enc_out;
        }
      }
      if (is_final) {                                 // next_states = None
        h.enc_state = false; h.feat_carry = 0; h.sub_n = 0; h.n_proc = 0;
        h.has_prev_addin = false; h.has_past_ctx = false;
      }
    }
    // ---------------- decode-block trigger (beam_search.py:590-634)
    while (h.enc_len > 0 && (kBlock - kLook) + kHopB * h.processed_block < h.enc_len) {
      p.dq_T.push_back((kBlock - kLook) + kHopB * h.processed_block);
      p.dq_final.push_back(0);
      h.processed_block++;
    }
    if (is_final && h.enc_len > 0) { p.dq_T.push_back(h.enc_len); p.dq_final.push_back(1); }
  }

  std::vector<StreamHost> hs_;
};

}  // namespace scb
