#include "fatcube_codec.h"
#include <cmath>
#include <cstring>
#include <dirent.h>
#include <sys/stat.h>

namespace pbso {
namespace {

struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    Reader(const uint8_t* d, size_t n) : p(d), end(d + n) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0; int shift = 0;
        while (p < end && shift < 70) {
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false; return 0;
    }
    double f64() {
        if (end - p < 8) { ok = false; return 0; }
        double d; std::memcpy(&d, p, 8); p += 8; return d;   // little-endian host
    }
    Reader sub() {
        uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Reader(p, 0); }
        Reader r(p, (size_t)n); p += n; return r;
    }
    void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: if (end - p < 8) ok = false; else p += 8; break;
            case 2: sub(); break;
            case 5: if (end - p < 4) ok = false; else p += 4; break;
            default: ok = false;
        }
    }
};

bool parse_vec(Reader r, std::vector<double>& v) {           // message vec { repeated double item = 1; }
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        if (f == 1 && wt == 2) { Reader s = r.sub(); while (!s.done() && s.ok) v.push_back(s.f64()); if (!s.ok) return false; }
        else if (f == 1 && wt == 1) v.push_back(r.f64());
        else r.skip(wt);
    }
    return r.ok;
}
bool parse_vec_i(Reader r, std::vector<int>& v) {            // message vec_i { repeated int32 item = 1; }
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        if (f == 1 && wt == 2) { Reader s = r.sub(); while (!s.done() && s.ok) v.push_back((int)(int64_t)s.varint()); if (!s.ok) return false; }
        else if (f == 1 && wt == 0) v.push_back((int)(int64_t)r.varint());
        else r.skip(wt);
    }
    return r.ok;
}
bool parse_mat(Reader r, std::vector<std::vector<double>>& m) {   // message mat { repeated vec item = 1; }
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        if (f == 1 && wt == 2) { m.emplace_back(); if (!parse_vec(r.sub(), m.back())) return false; }
        else r.skip(wt);
    }
    return r.ok;
}
bool parse_mat_i(Reader r, std::vector<std::vector<int>>& m) {
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        if (f == 1 && wt == 2) { m.emplace_back(); if (!parse_vec_i(r.sub(), m.back())) return false; }
        else r.skip(wt);
    }
    return r.ok;
}
bool parse_t1(Reader r, FatcubeMap& o) {                     // ffat_map.proto:30-38
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        bool good = true;
        if (f == 1 && wt == 1) o.cellsize = r.f64();
        else if (f == 2 && wt == 2) good = parse_mat(r.sub(), o.lowcorners);
        else if (f == 3 && wt == 2) good = parse_mat_i(r.sub(), o.n_elements);
        else if (f == 4 && wt == 2) good = parse_vec_i(r.sub(), o.strides);
        else if (f == 5 && wt == 2) good = parse_vec(r.sub(), o.center1);
        else if (f == 6 && wt == 2) good = parse_vec(r.sub(), o.bboxlow);
        else if (f == 7 && wt == 2) good = parse_vec(r.sub(), o.bboxtop);
        else r.skip(wt);
        if (!good) return false;
    }
    return r.ok;
}
bool parse_t3(Reader r, FatcubeMap& o) {                     // ffat_map.proto:40-47
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        bool good = true;
        if (f == 1 && wt == 1) o.k = r.f64();
        else if (f == 2 && wt == 2) good = parse_vec(r.sub(), o.center3);
        else if (f == 3 && wt == 2) good = parse_t1(r.sub(), o);
        else if (f == 4 && wt == 0) o.is_compressed = r.varint() != 0;
        else if (f == 5 && wt == 2) good = parse_mat(r.sub(), o.psi);
        else if (f == 6 && wt == 0) o.modeid = (int)(int64_t)r.varint();
        else r.skip(wt);
        if (!good) return false;
    }
    return r.ok;
}

void put_varint(std::string& s, uint64_t v) {
    while (v >= 0x80) { s.push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
    s.push_back((char)v);
}
void put_tag(std::string& s, int field, int wt) { put_varint(s, ((uint64_t)field << 3) | wt); }
void put_f64(std::string& s, double d) { char b[8]; std::memcpy(b, &d, 8); s.append(b, 8); }
void put_bytes(std::string& s, int field, const std::string& payload) {
    put_tag(s, field, 2); put_varint(s, payload.size()); s += payload;
}
std::string enc_vec(const std::vector<double>& v) {
    std::string s;
    if (!v.empty()) { put_tag(s, 1, 2); put_varint(s, v.size() * 8); for (double d : v) put_f64(s, d); }
    return s;
}
std::string enc_vec_i(const std::vector<int>& v) {
    std::string s;
    if (!v.empty()) {
        std::string p; for (int x : v) put_varint(p, (uint64_t)(int64_t)x);   // int32: sign-extended
        put_bytes(s, 1, p);
    }
    return s;
}
std::string enc_mat(const std::vector<std::vector<double>>& m) {
    std::string s; for (auto& v : m) put_bytes(s, 1, enc_vec(v)); return s;
}
std::string enc_mat_i(const std::vector<std::vector<int>>& m) {
    std::string s; for (auto& v : m) put_bytes(s, 1, enc_vec_i(v)); return s;
}
}  // namespace

bool fatcube_decode(const uint8_t* data, size_t size, FatcubeMap& out, std::string& err) {
    Reader r(data, size);
    while (!r.done() && r.ok) {
        uint64_t tag = r.varint(); int f = (int)(tag >> 3), wt = (int)(tag & 7);
        if (f == 1 && wt == 2) { if (!parse_t3(r.sub(), out)) { err = "malformed ffat_map_t_3"; return false; } }
        else r.skip(wt);
    }
    if (!r.ok) { err = "truncated or malformed protobuf stream"; return false; }
    return true;
}

void fatcube_encode(const FatcubeMap& m, std::string& out) {
    std::string t1;
    if (m.cellsize != 0) { put_tag(t1, 1, 1); put_f64(t1, m.cellsize); }
    put_bytes(t1, 2, enc_mat(m.lowcorners));
    put_bytes(t1, 3, enc_mat_i(m.n_elements));
    put_bytes(t1, 4, enc_vec_i(m.strides));
    put_bytes(t1, 5, enc_vec(m.center1));
    put_bytes(t1, 6, enc_vec(m.bboxlow));
    put_bytes(t1, 7, enc_vec(m.bboxtop));
    std::string t3;
    if (m.k != 0) { put_tag(t3, 1, 1); put_f64(t3, m.k); }
    put_bytes(t3, 2, enc_vec(m.center3));
    put_bytes(t3, 3, t1);
    if (m.is_compressed) { put_tag(t3, 4, 0); put_varint(t3, 1); }
    put_bytes(t3, 5, enc_mat(m.psi));
    if (m.modeid != 0) { put_tag(t3, 6, 0); put_varint(t3, (uint64_t)(int64_t)m.modeid); }
    out.clear();
    put_bytes(out, 1, t3);
}

bool list_dir_files(const char* dirname, std::vector<std::string>& names, const char* contains) {
    DIR* dir = opendir(dirname);
    if (!dir) return false;                                   // reference: perror(""), empty list
    struct dirent* ent;
    while ((ent = readdir(dir)) != nullptr) {
        std::string f = dirname + std::string("/") + std::string(ent->d_name);
        struct stat st;
        const bool is_file = stat(f.c_str(), &st) == 0;       // IsFile (io.cpp:36-44): "exists"
        if (is_file && ent->d_name[0] != '.' && contains && f.find(contains) != std::string::npos)
            names.push_back(f);
    }
    closedir(dir);
    return true;
}

}  // namespace pbso
