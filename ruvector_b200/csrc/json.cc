// serde-JSON weight wire format of RuvectorLayer (the only public ingress /
// egress of the reference's weights: all fields are private, the structs derive
// Serialize/Deserialize -- crates/ruvector-gnn/src/layer.rs:13,50,85,214,314;
// used by to_json/from_json in crates/ruvector-gnn-node/src/lib.rs:97-117).
//
//   {"w_msg":L,"w_agg":L,"w_update":{"w_z":L,"u_z":L,"w_r":L,"u_r":L,"w_h":L,"u_h":L},
//    "attention":{"num_heads":h,"head_dim":d,"q_linear":L,"k_linear":L,"v_linear":L,"out_linear":L},
//    "norm":{"gamma":A,"beta":A,"eps":1e-5},"dropout":p}
//   L = {"weights":{"v":1,"dim":[out,in],"data":[...]},"bias":{"v":1,"dim":[out],"data":[...]}}
//
// ({"v","dim","data"} is ndarray's serde form.)  Numbers are parsed with
// strtof so the f32 bit pattern serde_json printed (shortest round-trip) is
// recovered exactly; they are printed as the shortest decimal that strtof
// maps back to the same float.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "layer_host.h"

namespace ruvgnn {
namespace {

struct Parser {
  const char *p, *end;
  std::string err;

  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool fail(const std::string &m) {
    if (err.empty()) err = m + " at byte " + std::to_string((long long)(p - (end - total)));
    return false;
  }
  size_t total = 0;
  bool expect(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return fail(std::string("expected '") + c + "'");
  }
  bool peek(char c) {
    ws();
    return p < end && *p == c;
  }
  bool string(std::string &s) {
    ws();
    if (p >= end || *p != '"') return fail("expected string");
    ++p;
    s.clear();
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) ++p;
      s.push_back(*p++);
    }
    if (p >= end) return fail("unterminated string");
    ++p;
    return true;
  }
  bool number(float &f) {
    ws();
    if (p >= end) return fail("expected number");
    // strtof needs a NUL-terminated buffer; numbers are short, copy the token.
    char tok[64];
    size_t n = 0;
    auto numeric = [](char c) {
      return (c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E';
    };
    while (p + n < end && n < 63 && numeric(p[n])) ++n;
    if (n == 0) return fail("expected number");
    memcpy(tok, p, n);
    tok[n] = 0;
    char *e = nullptr;
    f = strtof(tok, &e);
    if (e == tok) return fail("bad number");
    p += (e - tok);
    return true;
  }
  bool float_array(std::vector<float> &v) {
    if (!expect('[')) return false;
    v.clear();
    if (peek(']')) {
      ++p;
      return true;
    }
    for (;;) {
      float f;
      if (!number(f)) return false;
      v.push_back(f);
      ws();
      if (p < end && *p == ',') {
        ++p;
        continue;
      }
      return expect(']');
    }
  }
  // skip any value.  Nesting is bounded like serde_json's recursion limit (128) so a hostile
  // document cannot overflow the stack.
  int depth = 0;
  struct DepthGuard {
    int &d;
    explicit DepthGuard(int &dd) : d(dd) { ++d; }
    ~DepthGuard() { --d; }
  };
  bool literal(const char *word) {
    const size_t n = strlen(word);
    if ((size_t)(end - p) < n || memcmp(p, word, n) != 0) return false;
    p += n;
    return true;
  }
  bool skip() {
    DepthGuard guard(depth);
    if (depth > 128) return fail("recursion limit exceeded");
    ws();
    if (p >= end) return fail("unexpected end");
    if (*p == '{') {
      ++p;
      if (peek('}')) {
        ++p;
        return true;
      }
      for (;;) {
        std::string k;
        if (!string(k) || !expect(':') || !skip()) return false;
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        return expect('}');
      }
    }
    if (*p == '[') {
      ++p;
      if (peek(']')) {
        ++p;
        return true;
      }
      for (;;) {
        if (!skip()) return false;
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        return expect(']');
      }
    }
    if (*p == '"') {
      std::string s;
      return string(s);
    }
    if (literal("true") || literal("false") || literal("null")) return true;
    float f;
    return number(f);
  }
  // a shape entry: a non-negative integer that fits u32 (anything else is rejected before a cast)
  static bool dim_ok(float d) { return d >= 0.f && d <= 4294967040.f && d == floorf(d); }

  // ndarray serde form {"v":1,"dim":[...],"data":[...]}
  bool ndarray(std::vector<float> &dim, std::vector<float> &data) {
    if (!expect('{')) return false;
    bool have_dim = false, have_data = false;
    for (;;) {
      std::string k;
      if (!string(k) || !expect(':')) return false;
      if (k == "dim") {
        ws();
        if (peek('[')) {
          if (!float_array(dim)) return false;
        } else {  // Ix1 may serialize its dim as a bare integer
          float f;
          if (!number(f)) return false;
          dim.assign(1, f);
        }
        have_dim = true;
      } else if (k == "data") {
        if (!float_array(data)) return false;
        have_data = true;
      } else if (k == "v") {
        float f;
        if (!number(f)) return false;
        if (f != 1.f) return fail("unknown ndarray format version");  // ndarray: "unknown array version"
      } else if (!skip()) {
        return false;
      }
      ws();
      if (p < end && *p == ',') {
        ++p;
        continue;
      }
      if (!expect('}')) return false;
      break;
    }
    if (!have_dim || !have_data) return fail("ndarray representation needs dim and data");
    size_t prod = 1;
    for (float d : dim) {
      if (!dim_ok(d)) return fail("invalid dimension (must be a non-negative integer below 2^32)");
      const size_t di = (size_t)d;
      if (di && prod > data.size() / di + 1) return fail("data and dimension must match in size");
      prod *= di;
    }
    if (prod != data.size()) return fail("data and dimension must match in size");
    return true;
  }
  bool linear(HostLinear &l) {
    if (!expect('{')) return false;
    bool hw = false, hb = false;
    for (;;) {
      std::string k;
      if (!string(k) || !expect(':')) return false;
      std::vector<float> dim;
      if (k == "weights") {
        if (!ndarray(dim, l.w)) return false;
        if (dim.size() != 2) return fail("weights must be 2-d");
        l.out = (uint32_t)dim[0];
        l.in = (uint32_t)dim[1];
        hw = true;
      } else if (k == "bias") {
        if (!ndarray(dim, l.b)) return false;
        if (dim.size() != 1) return fail("bias must be 1-d");
        hb = true;
      } else if (!skip()) {
        return false;
      }
      ws();
      if (p < end && *p == ',') {
        ++p;
        continue;
      }
      if (!expect('}')) return false;
      break;
    }
    if (!hw || !hb) return fail("Linear needs weights and bias");
    if (l.b.size() != l.out) return fail("bias length != weight rows");
    return true;
  }
};

void put_float(std::string &s, float f) {
  if (!isfinite(f)) {  // serde_json writes null for non-finite floats
    s += "null";
    return;
  }
  char buf[48];
  for (int prec = 1; prec <= 9; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
    if (strtof(buf, nullptr) == f) break;
  }
  if (!strpbrk(buf, ".eEn")) strcat(buf, ".0");
  s += buf;
}
void put_array(std::string &s, const std::vector<float> &v) {
  s += '[';
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) s += ',';
    put_float(s, v[i]);
  }
  s += ']';
}
void put_linear(std::string &s, const HostLinear &l) {
  s += "{\"weights\":{\"v\":1,\"dim\":[" + std::to_string(l.out) + "," + std::to_string(l.in) +
       "],\"data\":";
  put_array(s, l.w);
  s += "},\"bias\":{\"v\":1,\"dim\":[" + std::to_string(l.out) + "],\"data\":";
  put_array(s, l.b);
  s += "}}";
}

}  // namespace

std::string parse_layer_json(const char *json, size_t len, LayerHost &L) {
  Parser ps{json, json + len, std::string()};
  ps.total = len;
  bool seen[L_COUNT] = {false};
  bool have_norm = false, have_dropout = false, have_heads = false;
  uint32_t head_dim = 0;
  auto obj = [&](auto &&on_key) -> bool {
    if (!ps.expect('{')) return false;
    if (ps.peek('}')) {
      ++ps.p;
      return true;
    }
    for (;;) {
      std::string k;
      if (!ps.string(k) || !ps.expect(':')) return false;
      if (!on_key(k)) return false;
      ps.ws();
      if (ps.p < ps.end && *ps.p == ',') {
        ++ps.p;
        continue;
      }
      return ps.expect('}');
    }
  };
  auto lin = [&](int id) -> bool {
    seen[id] = true;
    return ps.linear(L.lin[id]);
  };
  bool ok = obj([&](const std::string &k) -> bool {
    if (k == "w_msg") return lin(L_W_MSG);
    if (k == "w_agg") return lin(L_W_AGG);
    if (k == "w_update")
      return obj([&](const std::string &g) -> bool {
        if (g == "w_z") return lin(L_W_Z);
        if (g == "u_z") return lin(L_U_Z);
        if (g == "w_r") return lin(L_W_R);
        if (g == "u_r") return lin(L_U_R);
        if (g == "w_h") return lin(L_W_H);
        if (g == "u_h") return lin(L_U_H);
        return ps.skip();
      });
    if (k == "attention")
      return obj([&](const std::string &g) -> bool {
        float f;
        if (g == "num_heads") {
          if (!ps.number(f)) return false;
          if (!Parser::dim_ok(f)) return ps.fail("num_heads must be a non-negative integer");
          L.heads = (uint32_t)f;
          have_heads = true;
          return true;
        }
        if (g == "head_dim") {
          if (!ps.number(f)) return false;
          if (!Parser::dim_ok(f)) return ps.fail("head_dim must be a non-negative integer");
          head_dim = (uint32_t)f;
          return true;
        }
        if (g == "q_linear") return lin(L_Q);
        if (g == "k_linear") return lin(L_K);
        if (g == "v_linear") return lin(L_V);
        if (g == "out_linear") return lin(L_OUT);
        return ps.skip();
      });
    if (k == "norm")
      return obj([&](const std::string &g) -> bool {
        std::vector<float> dim;
        if (g == "gamma") return ps.ndarray(dim, L.gamma);
        if (g == "beta") return ps.ndarray(dim, L.beta);
        if (g == "eps") {
          have_norm = true;
          return ps.number(L.eps);
        }
        return ps.skip();
      });
    if (k == "dropout") {
      have_dropout = true;
      return ps.number(L.dropout);
    }
    return ps.skip();
  });
  if (!ok) return "Deserialization error: " + ps.err;
  for (int i = 0; i < L_COUNT; ++i)
    if (!seen[i]) return "Deserialization error: missing field (linear #" + std::to_string(i) + ")";
  if (!have_norm || !have_dropout || !have_heads)
    return "Deserialization error: missing field (norm/dropout/num_heads)";
  L.input_dim = L.lin[L_W_MSG].in;
  L.hidden_dim = L.lin[L_W_MSG].out;
  const uint32_t H = L.hidden_dim;
  for (int i = 0; i < L_COUNT; ++i) {
    const uint32_t want_in = (i == L_W_MSG) ? L.input_dim : H;
    if (L.lin[i].out != H || L.lin[i].in != want_in)
      return "Deserialization error: linear #" + std::to_string(i) + " has inconsistent shape";
  }
  if (L.gamma.size() != H || L.beta.size() != H)
    return "Deserialization error: norm gamma/beta length != hidden_dim";
  if (L.heads == 0 || H % L.heads != 0 || (head_dim && head_dim != H / L.heads))
    return "Deserialization error: num_heads/head_dim inconsistent with hidden_dim";
  return "";
}

std::string emit_layer_json(const LayerHost &L) {
  std::string s;
  size_t n = 0;
  for (int i = 0; i < L_COUNT; ++i) n += L.lin[i].w.size() + L.lin[i].b.size();
  s.reserve(n * 14 + 1024);
  s += "{\"w_msg\":";
  put_linear(s, L.lin[L_W_MSG]);
  s += ",\"w_agg\":";
  put_linear(s, L.lin[L_W_AGG]);
  s += ",\"w_update\":{\"w_z\":";
  put_linear(s, L.lin[L_W_Z]);
  s += ",\"u_z\":";
  put_linear(s, L.lin[L_U_Z]);
  s += ",\"w_r\":";
  put_linear(s, L.lin[L_W_R]);
  s += ",\"u_r\":";
  put_linear(s, L.lin[L_U_R]);
  s += ",\"w_h\":";
  put_linear(s, L.lin[L_W_H]);
  s += ",\"u_h\":";
  put_linear(s, L.lin[L_U_H]);
  s += "},\"attention\":{\"num_heads\":" + std::to_string(L.heads) +
       ",\"head_dim\":" + std::to_string(L.heads ? L.hidden_dim / L.heads : 0) + ",\"q_linear\":";
  put_linear(s, L.lin[L_Q]);
  s += ",\"k_linear\":";
  put_linear(s, L.lin[L_K]);
  s += ",\"v_linear\":";
  put_linear(s, L.lin[L_V]);
  s += ",\"out_linear\":";
  put_linear(s, L.lin[L_OUT]);
  s += "},\"norm\":{\"gamma\":{\"v\":1,\"dim\":[" + std::to_string(L.gamma.size()) + "],\"data\":";
  put_array(s, L.gamma);
  s += "},\"beta\":{\"v\":1,\"dim\":[" + std::to_string(L.beta.size()) + "],\"data\":";
  put_array(s, L.beta);
  s += "},\"eps\":";
  put_float(s, L.eps);
  s += "},\"dropout\":";
  put_float(s, L.dropout);
  s += "}";
  return s;
}

}  // namespace ruvgnn
