// Workspace planning: one code path both sizes the caller's workspace and carves it.
#include <algorithm>

#include "common.cuh"
#include "loss.cuh"
#include "bf16_path.cuh"

namespace an3d {

int plan_f32(const Model& m, int B, int N, int flags, void* workspace, PlanF32* p) {
  if (B < 1 || N < 1) {
    set_error("batch (%d) and num_points (%d) must be >= 1", B, N);
    return AN3D_ERR_INVALID;
  }
  const bool training = (flags & AN3D_TRAINING) != 0;
  const int nprec = ((flags & AN3D_PRECISION_BF16) != 0) + ((flags & AN3D_PRECISION_BF16X3) != 0) + ((flags & AN3D_PRECISION_BF16X6) != 0);
  if (nprec > 1) {
    set_error("flags name more than one precision mode");
    return AN3D_ERR_INVALID;
  }
  // bf16: the fused conv-stack kernels where the architecture has their shape, else the materialised path with one
  // bf16 image per GEMM operand; bf16x3 / bf16x6: the materialised path with two / three images
  const bool bf16 = (flags & AN3D_PRECISION_BF16) != 0 && bf16_supported(m) == AN3D_OK;
  int tc_split = 0;
  if ((flags & AN3D_PRECISION_BF16) != 0 && !bf16) tc_split = 1;
  if (flags & AN3D_PRECISION_BF16X3) tc_split = 2;
  if (flags & AN3D_PRECISION_BF16X6) tc_split = 3;
  p->tc_split = tc_split;
  Arena a;
  a.base = static_cast<char*>(workspace);
  p->B = B;
  p->N = N;
  p->bf16 = bf16;
  p->M = (int64_t)B * N;
  const int64_t M = p->M;
  const int c_emb = m.conv[EMB].back().cout;
  p->feat = a.take<float>((int64_t)B * 2 * c_emb);
  int64_t max_conv_ch = 0, max_fc_ch = 0;
  for (int s = 0; s < 3; ++s) {
    for (int br = 0; br < 2; ++br) {
      p->pin[s][br] = bf16 ? nullptr : a.take<float>(M * 3);
      for (size_t l = 0; l < m.conv[s].size(); ++l) {
        p->z[s][l][br] = bf16 ? nullptr : a.take<float>(M * m.conv[s][l].cout);
        max_conv_ch = std::max<int64_t>(max_conv_ch, m.conv[s][l].cout);
      }
      const int c3 = m.conv[s].back().cout;
      if (s == EMB) {
        p->g[s][br] = p->feat + (int64_t)br * c_emb;  // interleaved [B, 2*C] head input
      } else {
        p->g[s][br] = a.take<float>((int64_t)B * c3);
      }
      p->gidx[s][br] = a.take<int32_t>((int64_t)B * c3);
    }
  }
  for (int s = 0; s < 3; ++s) {
    const int nbr = s == HEAD ? 1 : 2;
    for (int br = 0; br < 2; ++br)
      for (size_t l = 0; l < m.fc[s].size(); ++l) {
        p->fz[s][l][br] = br < nbr ? a.take<float>((int64_t)B * m.fc[s][l].cout) : nullptr;
        max_fc_ch = std::max<int64_t>(max_fc_ch, m.fc[s][l].cout);
        max_fc_ch = std::max<int64_t>(max_fc_ch, m.fc[s][l].cin);
      }
  }
  for (int s = 0; s < 3; ++s)
    for (int l = 0; l <= AN3D_MAX_LAYERS; ++l) {
      p->fcw[s][l] = nullptr;
      p->fcx[s][l][0] = p->fcx[s][l][1] = nullptr;
    }
  if (bf16) {
    for (int s = 0; s < 3; ++s) {
      const int nbr = s == HEAD ? 1 : 2;
      for (size_t l = 0; l < m.fc[s].size(); ++l) {
        const Lin& L = m.fc[s][l];
        p->fcw[s][l] = a.take<__nv_bfloat16>(fc_image_elems(L.cin, L.cout));
        for (int br = 0; br < nbr; ++br) p->fcx[s][l][br] = a.take<__nv_bfloat16>(fc_image_elems(B, L.cin));
      }
    }
  }
  // dropout masks: s1/b0, s1/b1, s2/b0, s2/b1, head; width = last hidden FC width of the stage
  for (int i = 0; i < 5; ++i) {
    const int s = i < 2 ? S1 : (i < 4 ? S2 : HEAD);
    const int width = m.fc[s][m.fc[s].size() - 2].cout;
    p->mask[i] = a.take<float>((int64_t)B * width);
  }
  for (int br = 0; br < 2; ++br) {
    p->mu[br] = a.take<float>((int64_t)B * 3);
    p->ang[br] = a.take<float>(B);
    p->angk[br] = a.take<int32_t>(B);
  }
  const int64_t nbn = m.bn_total_ch();
  p->bn.scale = a.take<float>(nbn);
  p->bn.shift = a.take<float>(nbn);
  p->bn.mean = a.take<float>(nbn);
  p->bn.inv = a.take<float>(nbn);
  p->bn.acc0 = a.take<double>(nbn);
  p->bn.acc1 = a.take<double>(nbn);
  p->loss_scratch = a.take<float>(loss_scratch_floats(B));
  p->dend = a.take<float>((int64_t)B * (5 * 3 + 3 * 2 * m.nb));
  if (training) {
    const int64_t big = bf16 ? 0 : M * max_conv_ch;
    p->dbuf[0] = a.take<float>(big);
    p->dbuf[1] = a.take<float>(big);
    p->dfc[0] = a.take<float>((int64_t)B * max_fc_ch);
    p->dfc[1] = a.take<float>((int64_t)B * max_fc_ch);
    p->dfeat = a.take<float>((int64_t)B * 2 * c_emb);
    p->dpin = a.take<float>(M * 3);
    int64_t max_c3 = 0;
    for (int s = 0; s < 3; ++s) max_c3 = std::max<int64_t>(max_c3, m.conv[s].back().cout);
    p->dg = a.take<float>((int64_t)B * max_c3);
    p->dout = a.take<float>((int64_t)B * (3 + 2 * m.nb));
    p->dbias_acc = a.take<double>(std::max<int64_t>(std::max(max_conv_ch, max_fc_ch), 3 + 2 * m.nb));
    if (bf16) {
      for (int br = 0; br < 2; ++br)
        p->fcdz[br] = a.take<__nv_bfloat16>(fc_image_elems(B, (int)std::max<int64_t>(max_fc_ch, 3 + 2 * m.nb)));
      p->dfc_b[0] = a.take<float>((int64_t)B * max_fc_ch);
      p->dfc_b[1] = a.take<float>((int64_t)B * max_fc_ch);
      p->dg_b = a.take<float>((int64_t)B * max_c3);
      p->dout_b = a.take<float>((int64_t)B * (3 + 2 * m.nb));
    }
    for (int br = 0; br < 2; ++br) {
      p->dc1[br] = a.take<float>((int64_t)B * 3);
      p->dc2[br] = a.take<float>((int64_t)B * 3);
      p->dang[br] = a.take<float>(B);
    }
  } else {
    p->dbuf[0] = p->dbuf[1] = p->dfc[0] = p->dfc[1] = p->dfeat = p->dpin = p->dg = p->dout = nullptr;
    p->dbias_acc = nullptr;
    for (int br = 0; br < 2; ++br) p->dc1[br] = p->dc2[br] = p->dang[br] = nullptr;
  }
  if (bf16) plan_bf16(m, B, N, flags, a, &p->bf);
  for (int s = 0; s < 3; ++s)
    for (int l = 0; l < AN3D_MAX_LAYERS; ++l) p->tcx[s][l][0] = p->tcx[s][l][1] = nullptr;
  if (tc_split > 0) {
    // image scratch of the split-operand GEMMs (gemm_tc.cuh), reused layer after layer: [0] the layer's input
    // activations, [1] the gradient at its output (training), [2] its weights.  Layers with a dimension below 8 stay on
    // the CUDA cores (the 3-wide first conv layer, 3-wide outputs).
    int64_t e0 = 0, e1 = 0, e2 = 0;
    auto lin = [&](const Lin& L, int64_t rows) {
      e0 = std::max(e0, fc_image_elems((int)rows, L.cin));
      e1 = std::max(e1, fc_image_elems((int)rows, L.cout));
      e2 = std::max(e2, fc_image_elems(L.cin, L.cout));
    };
    for (int s = 0; s < 3; ++s) {
      for (size_t l = 1; l < m.conv[s].size(); ++l) lin(m.conv[s][l], M);
      for (size_t l = 0; l < m.fc[s].size(); ++l) lin(m.fc[s][l], B);
    }
    p->tcbuf_elems[0] = e0 * tc_split;
    p->tcbuf_elems[1] = training ? e1 * tc_split : 0;
    p->tcbuf_elems[2] = e2 * tc_split;
    for (int i = 0; i < 3; ++i) p->tcbuf[i] = a.take<__nv_bfloat16>(p->tcbuf_elems[i]);
    p->bwd_coef = training ? a.take<float>(2 * ((max_conv_ch + 3) & ~int64_t(3))) : nullptr;
    // the conv layers' input images stay for the backward's wgrad (the pack applies BN + ReLU of the producing layer:
    // re-packing them read every activation a second time)
    if (training)
      for (int s = 0; s < 3; ++s)
        for (size_t l = 1; l < m.conv[s].size(); ++l) {
          const Lin& L = m.conv[s][l];
          if (std::min(std::min<int64_t>(L.cin, L.cout), M) < 8) continue;        // (tcg::use_tensor_cores)
          for (int br = 0; br < 2; ++br) p->tcx[s][l][br] = a.take<__nv_bfloat16>(fc_image_elems((int)M, L.cin) * tc_split);
        }
  }
  p->bytes = (a.off + 255) & ~int64_t(255);
  return AN3D_OK;
}

}  // namespace an3d
