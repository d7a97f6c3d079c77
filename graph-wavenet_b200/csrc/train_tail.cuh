// Step tail of engine.trainer.train (engine.py:46-58) as device kernels, so that a whole optimisation step is a
// fixed launch sequence with no host round trip:
//   loss_reduce / loss_grad : inverse_transform (Utils/util.py:116-117) + masked_mae / masked_mape / masked_rmse
//                             (Utils/util.py:510-552) and d(masked_mae)/d(output) written straight into the head's
//                             gradient operand;
//   gradnorm / adam         : clip_grad_norm_(params, clip) (engine.py:53-54) + torch.optim.Adam with L2 weight decay
//                             (engine.py:33,55) over the flat gradient / parameter buffers.
// Step counter, dropout seed and the reduction accumulators live in a small device control block so that the same
// launches can be replayed from a CUDA graph.
#pragma once
#include "elementwise.cuh"

namespace gwn {

struct TrainCtrl {
  unsigned long long seed;   // Philox key of the current step's dropout draws
  long long step;            // Adam step count
  double acc[6];             // 0 mask count, 1 sum|d|, 2 sum|d|/label, 3 sum d^2, 4 sum g^2, 5 spare
};

GWN_DEV void reduce_add_d(double* dst, double v) {
#if GWN_EMU
  *dst += v;
#else
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
#endif
}

GWN_GLOBAL train_begin_kernel(TrainCtrl* c) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, 1) {
    c->seed += 0x9E3779B97F4A7C15ull;
    for (int k = 0; k < 6; ++k) c->acc[k] = 0.0;
  }
}

// out: network output [B,O,N,T] contiguous; label(b,n,o) = y[b*ys0 + n*ys1 + o*ys2] broadcast over T (engine.py:46-48).
GWN_GLOBAL loss_reduce_kernel(const float* out, const float* y, i64 ys0, i64 ys1, i64 ys2, float mean, float std, int B, int O,
                              int N, int T, TrainCtrl* c) {
  GWN_PDL_ENTRY();
  double cnt = 0.0, sa = 0.0, sm = 0.0, sq = 0.0;
  GWN_FOR_EACH(i, (i64)B * O * N * T) {
    const int t = (int)(i % T);
    i64 r = i / T;
    const int n = (int)(r % N);
    r /= N;
    const int o = (int)(r % O);
    const i64 b = r / O;
    const float lab = y[b * ys0 + n * ys1 + o * ys2];
    if (lab != 0.0f) {   // mask = labels != null_val (0.0); NaN labels stay in the mask, their terms are zeroed below
      if (t == 0) cnt += 1.0;
      const float d = (out[i] * std + mean) - lab;
      const float ad = fabsf(d);
      if (ad == ad) {     // torch.where(isnan(loss), 0, loss)
        sa += ad;
        sq += (double)d * d;
        const float mp = ad / lab;
        if (mp == mp) sm += mp;
      }
    }
  }
#if GWN_EMU
  reduce_add_d(&c->acc[0], cnt);
  reduce_add_d(&c->acc[1], sa);
  reduce_add_d(&c->acc[2], sm);
  reduce_add_d(&c->acc[3], sq);
#else
  // block-level reduction, then ONE set of four fp64 atomics per block: a set per warp (5000 warps onto the same four
  // addresses) made this 160k-element pass take 28 us
  __shared__ double red[8][4];
  double v[4] = {warp_sum(cnt), warp_sum(sa), warp_sum(sm), warp_sum(sq)};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int y = 0; y < (int)(blockDim.x >> 5); ++y) t += red[y][threadIdx.x];
    if (t != 0.0) atomicAdd(&c->acc[threadIdx.x], t);
  }
#endif
}

// dout[((b*T+t)*N+n)*ldo + o] = d masked_mae / d out[b,o,n,t];  metrics = {mae, mape, rmse}.
// masked_mae = mean(|d| * mask / mean(mask)) = sum|d| / (count * T): count = unmasked labels, each broadcast over T.
GWN_GLOBAL loss_grad_kernel(const float* out, const float* y, i64 ys0, i64 ys1, i64 ys2, float mean, float std, int B, int O,
                            int N, int T, int ldo, const TrainCtrl* c, float* dout, float* metrics) {
  GWN_PDL_ENTRY();
  const double denom = c->acc[0] * (double)T;
  const float gs = denom > 0.0 ? (float)((double)std / denom) : 0.0f;
  GWN_FOR_EACH(i, (i64)B * T * N * ldo) {
    const int o = (int)(i % ldo);
    i64 r = i / ldo;
    const int n = (int)(r % N);
    r /= N;
    const int t = (int)(r % T);
    const i64 b = r / T;
    float g = 0.0f;
    if (o < O) {
      const float lab = y[b * ys0 + n * ys1 + o * ys2];
      if (lab != 0.0f) {
        const float d = (out[((b * O + o) * N + n) * T + t] * std + mean) - lab;
        if (d == d) g = d > 0.0f ? gs : (d < 0.0f ? -gs : 0.0f);
      }
    }
    dout[i] = g;
    if (i == 0) {
      metrics[0] = denom > 0.0 ? (float)(c->acc[1] / denom) : 0.0f;
      metrics[1] = denom > 0.0 ? (float)(c->acc[2] / denom) : 0.0f;
      metrics[2] = denom > 0.0 ? (float)sqrt(c->acc[3] / denom) : 0.0f;
    }
  }
}

// metrics = {mae, mape, rmse} from the accumulators alone (trainer.eval: no gradient wanted)
GWN_GLOBAL metrics_finalize_kernel(const TrainCtrl* c, int T, float* metrics) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, 1) {
    const double denom = c->acc[0] * (double)T;
    metrics[0] = denom > 0.0 ? (float)(c->acc[1] / denom) : 0.0f;
    metrics[1] = denom > 0.0 ? (float)(c->acc[2] / denom) : 0.0f;
    metrics[2] = denom > 0.0 ? (float)sqrt(c->acc[3] / denom) : 0.0f;
  }
}

// hyper: lr, beta1, beta2, eps, weight_decay, max_norm (<= 0: no clipping), grad_scale (1/world after an all-reduce sum)
struct AdamHyper {
  float lr, beta1, beta2, eps, wd, max_norm, gscale, pad;
};

// sum of squares of the live gradient elements (live4: one byte per 4 floats); also advances the step counter.
GWN_GLOBAL gradnorm_kernel(const float* g, const uint8_t* live4, i64 n4, TrainCtrl* c) {
  GWN_PDL_ENTRY();
  double s = 0.0;
  GWN_FOR_EACH(i, n4) {
    if (live4[i]) {
      const float4 v = ld4(g + 4 * i);
      s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (i == 0) c->step += 1;
  }
  reduce_add_d(&c->acc[4], s);
}

// gin (nullable): read the gradient from here instead of g (the all-reduced sum of the data-parallel step); g always
// receives the clipped gradient.
GWN_GLOBAL adam_kernel(float* p, float* g, float* m, float* v, const uint8_t* live4, i64 n4, const AdamHyper* hp,
                       const TrainCtrl* c, float* metrics, const float* gin) {
  GWN_PDL_ENTRY();
  const AdamHyper h = *hp;
  const float total_norm = (float)sqrt(c->acc[4]) * h.gscale;
  float coef = h.gscale;
  if (h.max_norm > 0.0f) coef *= fminf(h.max_norm / (total_norm + 1e-6f), 1.0f);   // clip_grad_norm_
  const double step = (double)c->step;
  const float bc1 = (float)(1.0 - pow((double)h.beta1, step));
  const float bc2s = (float)sqrt(1.0 - pow((double)h.beta2, step));
  const float step_size = h.lr / bc1;
  GWN_FOR_EACH(i, n4) {
    if (i == 0 && metrics) metrics[3] = total_norm;
    if (!live4[i]) continue;
    float4 pp = ld4(p + 4 * i), gg = ld4((gin ? gin : g) + 4 * i), mm = ld4(m + 4 * i), vv = ld4(v + 4 * i);
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gc = ga[k] * coef;
      ga[k] = gc;                                   // p.grad after the step holds the clipped gradient
      const float gw = fmaf(h.wd, pa[k], gc);       // L2 weight decay folded into the gradient
      ma[k] = ma[k] + (gw - ma[k]) * (1.0f - h.beta1);
      va[k] = va[k] * h.beta2 + (1.0f - h.beta2) * gw * gw;
      const float denom = sqrtf(va[k]) / bc2s + h.eps;
      pa[k] = pa[k] - step_size * (ma[k] / denom);
    }
    st4(p + 4 * i, pp);
    st4(g + 4 * i, gg);
    st4(m + 4 * i, mm);
    st4(v + 4 * i, vv);
  }
}

}  // namespace gwn
