// twr_safetensors.cpp -- native reader of a BasicPolicy / Conv1dPolicy checkpoint in safetensors format
// (what the reference's load_checkpoint / convert_pt_to_safetensors read and write, src/twisterl/utils.py:131-190),
// so that a C / C++ / Rust host can build a device policy without Python: the state-dict tensors are put into the
// layouts `to_rust()` hands to nn.Policy (src/twisterl/nn/utils.py:17-75) and passed to twr_policy_create.
//
// File format: u64 little-endian header length, a JSON object {"name": {"dtype": "F32", "shape": [..],
// "data_offsets": [begin, end]}, ..., "__metadata__": {..}}, then the raw little-endian tensor bytes.
#include "../../include/twisterl_b200.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <map>
#include <string>
#include <vector>

extern "C" int twr_set_error(int code, const char* msg);   // twr_engine.cu: records twr_last_error(), returns `code`

namespace {

struct Tensor { std::vector<int64_t> shape; std::vector<float> data; };

// Minimal JSON walk for the safetensors header (flat object of objects; strings without escapes we care about).
struct Parser {
    const std::string& s;
    size_t i = 0;
    explicit Parser(const std::string& str) : s(str) {}
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    bool str(std::string* out) {
        ws();
        if (i >= s.size() || s[i] != '"') return false;
        ++i;
        out->clear();
        while (i < s.size() && s[i] != '"') { if (s[i] == '\\' && i + 1 < s.size()) ++i; out->push_back(s[i++]); }
        if (i >= s.size()) return false;
        ++i;
        return true;
    }
    bool num(int64_t* out) {
        ws();
        size_t j = i;
        while (j < s.size() && (s[j] == '-' || (s[j] >= '0' && s[j] <= '9'))) ++j;
        if (j == i) return false;
        *out = std::strtoll(s.substr(i, j - i).c_str(), nullptr, 10);
        i = j;
        return true;
    }
    bool skip_value() {              // any JSON value (used for __metadata__ and unknown fields)
        ws();
        if (i >= s.size()) return false;
        if (s[i] == '"') { std::string t; return str(&t); }
        if (s[i] == '{' || s[i] == '[') {
            const char open = s[i], close = open == '{' ? '}' : ']';
            int depth = 0;
            bool in_str = false;
            for (; i < s.size(); ++i) {
                if (in_str) { if (s[i] == '\\') ++i; else if (s[i] == '"') in_str = false; continue; }
                if (s[i] == '"') in_str = true;
                else if (s[i] == open) ++depth;
                else if (s[i] == close && --depth == 0) { ++i; return true; }
            }
            return false;
        }
        while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']') ++i;
        return true;
    }
};

struct Entry { std::string dtype; std::vector<int64_t> shape; int64_t begin = 0, end = 0; };

bool parse_header(const std::string& js, std::map<std::string, Entry>* out) {
    Parser p(js);
    if (!p.eat('{')) return false;
    if (p.eat('}')) return true;
    do {
        std::string name;
        if (!p.str(&name) || !p.eat(':')) return false;
        if (name == "__metadata__") { if (!p.skip_value()) return false; continue; }
        if (!p.eat('{')) return false;
        Entry e;
        do {
            std::string key;
            if (!p.str(&key) || !p.eat(':')) return false;
            if (key == "dtype") { if (!p.str(&e.dtype)) return false; }
            else if (key == "shape" || key == "data_offsets") {
                std::vector<int64_t> v;
                if (!p.eat('[')) return false;
                if (!p.eat(']')) {
                    do { int64_t x; if (!p.num(&x)) return false; v.push_back(x); } while (p.eat(','));
                    if (!p.eat(']')) return false;
                }
                if (key == "shape") e.shape = v;
                else { if (v.size() != 2) return false; e.begin = v[0]; e.end = v[1]; }
            } else if (!p.skip_value()) return false;
        } while (p.eat(','));
        if (!p.eat('}')) return false;
        (*out)[name] = e;
    } while (p.eat(','));
    return p.eat('}');
}

int load_file(const char* path, std::map<std::string, Tensor>* out) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return twr_set_error(TWR_ERR_INVALID, (std::string("cannot open ") + path).c_str());
    uint64_t hlen = 0;
    if (std::fread(&hlen, 8, 1, f) != 1 || hlen == 0 || hlen > (1ull << 26)) { std::fclose(f); return twr_set_error(TWR_ERR_INVALID, "not a safetensors file (bad header length)"); }
    std::string js((size_t)hlen, '\0');
    if (std::fread(&js[0], 1, (size_t)hlen, f) != (size_t)hlen) { std::fclose(f); return twr_set_error(TWR_ERR_INVALID, "truncated safetensors header"); }
    std::map<std::string, Entry> hdr;
    if (!parse_header(js, &hdr)) { std::fclose(f); return twr_set_error(TWR_ERR_INVALID, "malformed safetensors header"); }
    const long data0 = 8 + (long)hlen;
    // the header is file content, not trusted input: every dimension, product and offset is range-checked before it sizes
    // an allocation or a read
    std::fseek(f, 0, SEEK_END);
    const int64_t data_bytes = (int64_t)std::ftell(f) - data0;
    for (const auto& kv : hdr) {
        const Entry& e = kv.second;
        if (e.dtype != "F32") { std::fclose(f); return twr_set_error(TWR_ERR_UNSUPPORTED, ("tensor " + kv.first + " is " + e.dtype + ": only F32 checkpoints are read").c_str()); }
        int64_t n = 1;
        bool ok = e.shape.size() <= 8;
        for (int64_t d : e.shape) {
            if (d < 0 || d >= (1ll << 31) || (d > 0 && n > (1ll << 40) / d)) { ok = false; break; }
            n *= d;
        }
        if (!ok || n > (1ll << 32)) { std::fclose(f); return twr_set_error(TWR_ERR_INVALID, ("tensor " + kv.first + ": shape out of range").c_str()); }
        if (e.begin < 0 || e.end < e.begin || e.end > data_bytes || e.end - e.begin != n * 4) {
            std::fclose(f);
            return twr_set_error(TWR_ERR_INVALID, ("tensor " + kv.first + ": data_offsets do not match the shape / file size").c_str());
        }
        Tensor t;
        t.shape = e.shape;
        t.data.resize((size_t)n);
        if (std::fseek(f, data0 + (long)e.begin, SEEK_SET) != 0 || (n > 0 && std::fread(t.data.data(), 4, (size_t)n, f) != (size_t)n)) {
            std::fclose(f);
            return twr_set_error(TWR_ERR_INVALID, ("tensor " + kv.first + ": truncated data").c_str());
        }
        (*out)[kv.first] = std::move(t);
    }
    std::fclose(f);
    return TWR_OK;
}

// torch Linear weight [out][in] -> nn.Linear's weights_vector W.T.flatten(): w[i*out + o]
std::vector<float> transpose(const Tensor& t) {
    const int64_t out = t.shape[0], in = t.shape[1];
    std::vector<float> w((size_t)(out * in));
    for (int64_t o = 0; o < out; ++o)
        for (int64_t i = 0; i < in; ++i) w[(size_t)(i * out + o)] = t.data[(size_t)(o * in + i)];
    return w;
}

}  // namespace

static int create_from_safetensors(twr_engine* e, const char* path, const int32_t* obs_shape, int32_t obs_shape_len,
                                   int32_t conv_dim, const int32_t* obs_perms, const int32_t* act_perms, int32_t n_perms,
                                   twr_policy** out);

extern "C" int twr_policy_create_from_safetensors(twr_engine* e, const char* path, const int32_t* obs_shape, int32_t obs_shape_len,
                                                  int32_t conv_dim, const int32_t* obs_perms, const int32_t* act_perms, int32_t n_perms,
                                                  twr_policy** out) {
    if (!e || !path || !out) return twr_set_error(TWR_ERR_INVALID, "NULL argument");
    try {                                                 // no C++ exception may cross the C boundary
        return create_from_safetensors(e, path, obs_shape, obs_shape_len, conv_dim, obs_perms, act_perms, n_perms, out);
    } catch (const std::exception& ex) {
        return twr_set_error(TWR_ERR_INVALID, (std::string("safetensors checkpoint: ") + ex.what()).c_str());
    } catch (...) {
        return twr_set_error(TWR_ERR_INVALID, "safetensors checkpoint: unexpected error");
    }
}

static int create_from_safetensors(twr_engine* e, const char* path, const int32_t* obs_shape, int32_t obs_shape_len,
                                   int32_t conv_dim, const int32_t* obs_perms, const int32_t* act_perms, int32_t n_perms,
                                   twr_policy** out) {
    std::map<std::string, Tensor> sd;
    int rc = load_file(path, &sd);
    if (rc) return rc;
    auto has = [&](const std::string& k) { return sd.count(k) != 0; };
    struct Stack { std::vector<std::vector<float>> w, b; std::vector<twr_linear_desc> d; };
    // Sequential of Linears and ReLUs: torch indices 0, 2, 4, ...; a Linear carries ReLU when another layer follows it
    // inside the stack or (common) always -- sequential_to_rust, nn/utils.py:17-44 with make_sequential's final_relu
    auto stack = [&](const std::string& name, bool relu_last, Stack* s) -> int {
        std::vector<int> idx;
        for (int k = 0; k < 64; ++k) if (has(name + "." + std::to_string(k) + ".weight")) idx.push_back(k);
        for (size_t n = 0; n < idx.size(); ++n) {
            const Tensor& W = sd[name + "." + std::to_string(idx[n]) + ".weight"];
            const std::string bk = name + "." + std::to_string(idx[n]) + ".bias";
            if (W.shape.size() != 2 || !has(bk) || (int64_t)sd[bk].data.size() != W.shape[0])
                return twr_set_error(TWR_ERR_INVALID, ("checkpoint: bad Linear " + name + "." + std::to_string(idx[n])).c_str());
            s->w.push_back(transpose(W));
            s->b.push_back(sd[bk].data);
        }
        for (size_t n = 0; n < idx.size(); ++n) {
            const Tensor& W = sd[name + "." + std::to_string(idx[n]) + ".weight"];
            s->d.push_back(twr_linear_desc{s->w[n].data(), s->b[n].data(), (int32_t)W.shape[1], (int32_t)W.shape[0],
                                           (relu_last || n + 1 < idx.size()) ? 1 : 0});
        }
        return TWR_OK;
    };
    Stack common, action, value;
    if ((rc = stack("common", true, &common)) || (rc = stack("action", false, &action)) || (rc = stack("value", false, &value))) return rc;

    twr_policy_desc d{};
    std::vector<float> vectors, bias;
    if (has("conv_layer.weight")) {                       // Conv1dPolicy: embeddingbag_to_rust's Conv1d branch (nn/utils.py:68-75)
        const Tensor& W = sd["conv_layer.weight"];        // [v][n_in][1]
        if (W.shape.size() != 3 || W.shape[2] != 1 || obs_shape_len != 2 || !obs_shape || (conv_dim != 0 && conv_dim != 1))
            return twr_set_error(TWR_ERR_INVALID, "Conv1d checkpoint needs a 2-D obs_shape and conv_dim 0 or 1");
        const int64_t v = W.shape[0], n_in = W.shape[1];
        if (obs_shape[0] < 1 || obs_shape[1] < 1 || obs_shape[0] > 65535 || obs_shape[1] > 65535 || v < 1 || n_in < 1 || v * obs_shape[1 - conv_dim] > (1 << 20))
            return twr_set_error(TWR_ERR_INVALID, "Conv1d checkpoint: obs_shape / kernel sizes out of range");
        vectors.resize((size_t)(v * n_in));
        for (int64_t r = 0; r < n_in; ++r)
            for (int64_t k = 0; k < v; ++k) vectors[(size_t)(r * v + k)] = W.data[(size_t)(k * n_in + r)];
        bias.assign((size_t)(v * obs_shape[1 - conv_dim]), 0.0f);
        d.obs_size = (int32_t)n_in; d.emb_size = (int32_t)bias.size();
        d.obs_shape[0] = obs_shape[0]; d.obs_shape[1] = obs_shape[1]; d.obs_shape_len = 2; d.conv_dim = conv_dim;
    } else {
        if (!has("embeddings.weight") || sd["embeddings.weight"].shape.size() != 2)
            return twr_set_error(TWR_ERR_INVALID, "checkpoint has no embeddings.weight / conv_layer.weight");
        const Tensor& W = sd["embeddings.weight"];        // torch Linear [E][obs_size] -> vec_vectors[obs][E] = W.T
        vectors = transpose(W);
        bias = has("embeddings.bias") ? sd["embeddings.bias"].data : std::vector<float>((size_t)W.shape[0], 0.0f);
        if ((int64_t)bias.size() != W.shape[0] || W.shape[0] < 1 || W.shape[1] < 1 || W.shape[1] >= 65536)
            return twr_set_error(TWR_ERR_INVALID, "checkpoint: embeddings.weight / embeddings.bias sizes do not match");
        d.obs_size = (int32_t)W.shape[1]; d.emb_size = (int32_t)W.shape[0];
        d.obs_shape[0] = d.obs_size; d.obs_shape_len = 1; d.conv_dim = 0;
    }
    d.emb_vectors = vectors.data(); d.emb_bias = bias.data(); d.emb_apply_relu = 1;
    d.common = common.d.data(); d.n_common = (int32_t)common.d.size();
    d.action_net = action.d.data(); d.n_action = (int32_t)action.d.size();
    d.value_net = value.d.data(); d.n_value = (int32_t)value.d.size();
    d.obs_perms = obs_perms; d.act_perms = act_perms; d.n_perms = n_perms;
    return twr_policy_create(e, &d, out);
}
