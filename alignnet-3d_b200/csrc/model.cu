// Host-only: architecture validation, flat parameter layout (TF variable names), error plumbing.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace an3d {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;

struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs: begin, end
  std::vector<int> tag;
};
static ProfState g_prof;

bool prof_active() { return g_prof.on; }

void prof_mark(int tag, bool begin, cudaStream_t st) {
  if (!g_prof.on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_prof.ev.push_back(e);
  if (begin) g_prof.tag.push_back(tag);
}
void prof_begin() {
  for (auto e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.tag.clear();
  g_prof.on = true;
}
void prof_end(float* ms, int* count) {
  g_prof.on = false;
  for (int i = 0; i < PROF_NTAGS; ++i) { ms[i] = 0.f; count[i] = 0; }
  for (size_t i = 0; i + 1 < g_prof.ev.size(); i += 2) {
    float t = 0.f;
    cudaEventSynchronize(g_prof.ev[i + 1]);
    if (cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]) == cudaSuccess) {
      const int tag = g_prof.tag[i / 2];
      ms[tag] += t;
      count[tag] += 1;
    }
  }
  for (auto e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.tag.clear();
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libalignnet_b200 has no CPU fallback", cudaGetErrorString(e));
    return AN3D_ERR_NO_DEVICE;
  }
  static thread_local int checked_dev = -1;
  if (checked_dev == dev) return AN3D_OK;
  int major = 0, minor = 0;
  AN3D_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  AN3D_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("device %d is sm_%d%d; libalignnet_b200 is built for sm_100a (B200) only", dev, major, minor);
    return AN3D_ERR_ARCH;
  }
  checked_dev = dev;
  return AN3D_OK;
}

static void add_tensor(std::vector<TensorInfo>& v, const std::string& name, int64_t off, std::initializer_list<int64_t> shp) {
  TensorInfo t;
  t.name = name;
  t.offset = off;
  t.ndim = (int)shp.size();
  int i = 0;
  for (auto s : shp) t.shape[i++] = s;
  for (; i < 4; ++i) t.shape[i] = 1;
  v.push_back(t);
}

// Scope strings follow the reference graph: models/tp8.py:52,62-66,77,92,140-143,154.
int build_model(const an3d_arch* a, Model* m) {
  if (!a || !m) {
    set_error("an3d_create: NULL argument");
    return AN3D_ERR_INVALID;
  }
  m->arch = *a;
  m->nb = a->num_bins;
  if (a->num_bins < 2 || a->num_bins > 256) {
    set_error("num_bins=%d out of range [2,256]", a->num_bins);
    return AN3D_ERR_INVALID;
  }
  for (int s = 0; s < 3; ++s) {
    if (a->n_conv[s] < 1 || a->n_conv[s] > AN3D_MAX_LAYERS || a->n_fc[s] < 1 || a->n_fc[s] > AN3D_MAX_LAYERS) {
      set_error("stage %d: need 1..%d conv layers and 1..%d hidden FC layers", s, AN3D_MAX_LAYERS, AN3D_MAX_LAYERS);
      return AN3D_ERR_INVALID;
    }
    for (int i = 0; i < a->n_conv[s]; ++i)
      if (a->conv[s][i] < 1 || a->conv[s][i] > 4096) {
        set_error("stage %d conv width %d invalid", s, a->conv[s][i]);
        return AN3D_ERR_INVALID;
      }
    for (int i = 0; i < a->n_fc[s]; ++i)
      if (a->fc[s][i] < 1 || a->fc[s][i] > 4096) {
        set_error("stage %d fc width %d invalid", s, a->fc[s][i]);
        return AN3D_ERR_INVALID;
      }
    if (!(a->keep_prob[s] > 0.f && a->keep_prob[s] <= 1.f)) {
      set_error("stage %d keep_prob %f must be in (0,1]", s, a->keep_prob[s]);
      return AN3D_ERR_INVALID;
    }
  }
  const char* conv_scope[3] = {"transformer1/embedding", "transformer2/embedding", "embedding"};
  const char* fc_scope[3] = {"transformer1/mlp/", "transformer2/mlp/", ""};
  const int out_dim[3] = {3, 3 + 2 * a->num_bins, 3 + 2 * a->num_bins};

  int64_t off = 0;
  m->bn_branch.clear();
  m->bn_head.clear();
  m->bn_branch_ch = m->bn_head_ch = 0;
  auto add_bn = [&](bool head, int ch, const std::string& scope) {
    BnLayer b;
    b.ch = ch;
    b.scope = scope;
    if (head) {
      b.choff = m->bn_head_ch;
      m->bn_head_ch += ch;
      m->bn_head.push_back(b);
      return (int)m->bn_head.size() - 1;
    }
    b.choff = m->bn_branch_ch;
    m->bn_branch_ch += ch;
    m->bn_branch.push_back(b);
    return (int)m->bn_branch.size() - 1;
  };
  auto add_lin = [&](std::vector<Lin>& dst, const std::string& prefix, const std::string& scope, int cin, int cout,
                     bool bn, bool head, bool first_conv) {
    Lin l;
    l.cin = cin;
    l.cout = cout;
    l.scope = scope;
    off = (off + 3) & ~int64_t(3);   // every tensor starts 16-byte aligned (vector loads / reductions)
    l.w = off;
    if (first_conv)
      add_tensor(m->trainable, prefix + scope + "/weights", off, {1, 3, 1, cout});
    else
      add_tensor(m->trainable, prefix + scope + "/weights", off, {cin, cout});
    off += (int64_t)cin * cout;
    off = (off + 3) & ~int64_t(3);
    l.b = off;
    add_tensor(m->trainable, prefix + scope + "/biases", off, {cout});
    off += cout;
    l.bn = bn ? add_bn(head, cout, scope) : -1;
    dst.push_back(l);
  };
  m->trainable.clear();
  m->state.clear();
  // shared weights, stage by stage: s1 conv, s1 fc, s2 conv, s2 fc, emb conv, head fc
  for (int s = 0; s < 3; ++s) {
    m->conv[s].clear();
    m->fc[s].clear();
  }
  for (int s = 0; s < 3; ++s) {
    int cin = 3;
    for (int i = 0; i < a->n_conv[s]; ++i) {
      add_lin(m->conv[s], "siamese/", std::string(conv_scope[s]) + "/conv" + std::to_string(i + 1), cin, a->conv[s][i],
              true, false, i == 0);
      cin = a->conv[s][i];
    }
    if (s < 2) {
      int fin = cin;
      for (int i = 0; i < a->n_fc[s]; ++i) {
        add_lin(m->fc[s], "siamese/", std::string(fc_scope[s]) + "fc" + std::to_string(i + 1), fin, a->fc[s][i], true,
                false, false);
        fin = a->fc[s][i];
      }
      add_lin(m->fc[s], "siamese/", std::string(fc_scope[s]) + "fc" + std::to_string(a->n_fc[s] + 1), fin, out_dim[s],
              false, false, false);
    }
  }
  {
    int fin = 2 * a->conv[2][a->n_conv[2] - 1];
    for (int i = 0; i < a->n_fc[2]; ++i) {
      add_lin(m->fc[2], "", "fc" + std::to_string(i + 1), fin, a->fc[2][i], true, true, false);
      fin = a->fc[2][i];
    }
    add_lin(m->fc[2], "", "fc" + std::to_string(a->n_fc[2] + 1), fin, out_dim[2], false, true, false);
  }
  // BN gamma/beta: branch 0 ("siamese/"), branch 1 ("siamese_1/", quirk Q0), head
  off = (off + 3) & ~int64_t(3);
  m->bn_base = off;
  int64_t soff = 0;
  for (int br = 0; br < 2; ++br) {
    const std::string prefix = br == 0 ? "siamese/" : "siamese_1/";
    for (auto& b : m->bn_branch) {
      add_tensor(m->trainable, prefix + b.scope + "/bn/gamma", off, {b.ch});
      off += b.ch;
      add_tensor(m->trainable, prefix + b.scope + "/bn/beta", off, {b.ch});
      off += b.ch;
      add_tensor(m->state, prefix + b.scope + "/bn/moments/Squeeze/ExponentialMovingAverage", soff, {b.ch});
      soff += b.ch;
      add_tensor(m->state, prefix + b.scope + "/bn/moments/Squeeze_1/ExponentialMovingAverage", soff, {b.ch});
      soff += b.ch;
    }
  }
  for (auto& b : m->bn_head) {
    add_tensor(m->trainable, b.scope + "/bn/gamma", off, {b.ch});
    off += b.ch;
    add_tensor(m->trainable, b.scope + "/bn/beta", off, {b.ch});
    off += b.ch;
    add_tensor(m->state, b.scope + "/bn/moments/Squeeze/ExponentialMovingAverage", soff, {b.ch});
    soff += b.ch;
    add_tensor(m->state, b.scope + "/bn/moments/Squeeze_1/ExponentialMovingAverage", soff, {b.ch});
    soff += b.ch;
  }
  m->n_trainable = (off + 3) & ~int64_t(3);
  m->n_state = soff;
  return AN3D_OK;
}

}  // namespace an3d

namespace an3d {
const char* last_error();
void prof_begin();
void prof_end(float* ms, int* count);
}

extern "C" {

int an3d_version(void) { return AN3D_VERSION; }

uint64_t an3d_launch_count(void) { return an3d::g_launch_count; }

int an3d_profile_begin(void) {
  an3d::prof_begin();
  return AN3D_OK;
}

int an3d_profile_end(float* ms_by_tag, int32_t* launches_by_tag) {
  if (!ms_by_tag || !launches_by_tag) {
    an3d::set_error("an3d_profile_end: NULL argument");
    return AN3D_ERR_INVALID;
  }
  an3d::prof_end(ms_by_tag, launches_by_tag);
  return AN3D_OK;
}

const char* an3d_last_error(void) { return an3d::last_error(); }

int an3d_create(const an3d_arch* arch, an3d_ctx** out_ctx) {
  if (!out_ctx) {
    an3d::set_error("an3d_create: out_ctx is NULL");
    return AN3D_ERR_INVALID;
  }
  an3d_ctx* c = new an3d_ctx();
  int r = an3d::build_model(arch, &c->impl.model);
  if (r != AN3D_OK) {
    delete c;
    *out_ctx = nullptr;
    return r;
  }
  *out_ctx = c;
  return AN3D_OK;
}

int an3d_destroy(an3d_ctx* ctx) {
  delete ctx;
  return AN3D_OK;
}

int an3d_num_elements(const an3d_ctx* ctx, int which, int64_t* out_count) {
  if (!ctx || !out_count || (which != 0 && which != 1)) {
    an3d::set_error("an3d_num_elements: bad argument");
    return AN3D_ERR_INVALID;
  }
  *out_count = which == 0 ? ctx->impl.model.n_trainable : ctx->impl.model.n_state;
  return AN3D_OK;
}

int an3d_num_tensors(const an3d_ctx* ctx, int which, int32_t* out_count) {
  if (!ctx || !out_count || (which != 0 && which != 1)) {
    an3d::set_error("an3d_num_tensors: bad argument");
    return AN3D_ERR_INVALID;
  }
  *out_count = (int32_t)(which == 0 ? ctx->impl.model.trainable.size() : ctx->impl.model.state.size());
  return AN3D_OK;
}

int an3d_tensor_info(const an3d_ctx* ctx, int which, int32_t index, char* name, int32_t name_capacity,
                     int64_t* offset, int32_t* ndim, int64_t shape[4]) {
  if (!ctx || (which != 0 && which != 1)) {
    an3d::set_error("an3d_tensor_info: bad argument");
    return AN3D_ERR_INVALID;
  }
  const auto& v = which == 0 ? ctx->impl.model.trainable : ctx->impl.model.state;
  if (index < 0 || index >= (int32_t)v.size()) {
    an3d::set_error("an3d_tensor_info: index %d out of range", index);
    return AN3D_ERR_INVALID;
  }
  const auto& t = v[index];
  if (name && name_capacity > 0) {
    strncpy(name, t.name.c_str(), name_capacity - 1);
    name[name_capacity - 1] = 0;
  }
  if (offset) *offset = t.offset;
  if (ndim) *ndim = t.ndim;
  if (shape)
    for (int i = 0; i < 4; ++i) shape[i] = t.shape[i];
  return AN3D_OK;
}

}  // extern "C"
