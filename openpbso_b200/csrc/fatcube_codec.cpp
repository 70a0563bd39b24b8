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

// ---------------------------------------------------------------------------------------------
// legacy igl::serialize reader / writer (format notes in fatcube_codec.h)
// ---------------------------------------------------------------------------------------------
namespace {
struct Cur {
    const uint8_t* p = nullptr; const uint8_t* end = nullptr; bool ok = true;
    Cur() {}
    Cur(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    bool need(size_t n) { if (!ok || (size_t)(end - p) < n) { ok = false; return false; } return true; }
    uint64_t u64() { uint64_t v = 0; if (need(8)) { std::memcpy(&v, p, 8); p += 8; } return v; }
    int32_t i32() { int32_t v = 0; if (need(4)) { std::memcpy(&v, p, 4); p += 4; } return v; }
    double f64() { double v = 0; if (need(8)) { std::memcpy(&v, p, 8); p += 8; } return v; }
    std::string str() { const uint64_t n = u64(); std::string s; if (need(n)) { s.assign((const char*)p, (size_t)n); p += n; } return s; }
    Cur sub(uint64_t n) { if (!need(n)) return Cur(end, end); Cur c(p, p + n); p += n; return c; }
};
struct Chunk { std::string name; Cur data; };
// next chunk of a member stream; false at the end (or on a malformed header: c.ok is cleared)
bool next_chunk(Cur& c, Chunk& out) {
    if (!c.ok || c.p >= c.end) return false;
    out.name = c.str();
    (void)c.str();                                                   // type: the writer's typeid name, not needed
    const uint64_t size = c.u64();
    out.data = c.sub(size);
    return c.ok;
}
bool read_matrix(Cur& c, std::vector<double>& v, int64_t& rows, int64_t& cols) {
    rows = (int64_t)c.u64(); cols = (int64_t)c.u64();
    if (!c.ok || rows < 0 || cols < 0 || (rows && cols > (int64_t)((size_t)(c.end - c.p) / 8) / rows)) { c.ok = false; return false; }
    v.resize((size_t)(rows * cols));
    for (auto& x : v) x = c.f64();
    return c.ok;
}
bool read_vec3(Cur& c, std::vector<double>& v) { int64_t r, q; return read_matrix(c, v, r, q) && v.size() == 3; }

void lg_put_u64(std::string& o, uint64_t v) { o.append((const char*)&v, 8); }
void lg_put_i32(std::string& o, int32_t v) { o.append((const char*)&v, 4); }
void lg_put_f64(std::string& o, double v) { o.append((const char*)&v, 8); }
void lg_put_str(std::string& o, const std::string& s) { lg_put_u64(o, s.size()); o += s; }
void lg_put_chunk(std::string& o, const char* name, const char* type, const std::string& data) {
    lg_put_str(o, name); lg_put_str(o, type); lg_put_u64(o, data.size()); o += data;
}
std::string mat_bytes(const std::vector<double>& v, int64_t rows, int64_t cols) {
    std::string o; lg_put_u64(o, (uint64_t)rows); lg_put_u64(o, (uint64_t)cols);
    for (double x : v) lg_put_f64(o, x);
    return o;
}
const char* T_INT = "i"; const char* T_DBL = "d"; const char* T_BOOL = "b";
const char* T_V3 = "N5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEE";
const char* T_MX = "N5Eigen6MatrixIdLin1ELin1ELi0ELin1ELin1EEE";
}  // namespace

bool legacy_fatcube_sniff(const uint8_t* data, size_t size) {
    static const char tag[] = "serial_map_ch3";
    uint64_t n = 0;
    if (size < 8 + sizeof(tag) - 1) return false;
    std::memcpy(&n, data, 8);
    return n == sizeof(tag) - 1 && std::memcmp(data + 8, tag, sizeof(tag) - 1) == 0;
}

bool legacy_fatcube_decode(const uint8_t* data, size_t size, FatcubeMap& out, std::string& err) {
    Cur file(data, data + size);
    Chunk top; bool found = false; Cur body(data, data);
    while (next_chunk(file, top))
        if (top.name == "serial_map_ch3") { body = top.data; found = true; }      // igl::deserialize keeps the LAST match (serialize.h:540-545)
    if (!file.ok) { err = "legacy .fatcube: truncated chunk header"; return false; }
    if (!found) { err = "legacy .fatcube: no \"serial_map_ch3\" object"; return false; }
    Cur members = body.sub(body.u64());
    if (!body.ok) { err = "legacy .fatcube: truncated object"; return false; }
    out = FatcubeMap();
    std::vector<double> psi, cpsi; int64_t pr = 0, pc = 0, cr = 0, cc = 0;
    bool have_shell = false;
    Chunk m;
    while (next_chunk(members, m)) {
        Cur& d = m.data;
        if (m.name == "modeId") out.modeid = d.i32();
        else if (m.name == "k") out.k = d.f64();
        else if (m.name == "center") read_vec3(d, out.center3);
        else if (m.name == "Psi") read_matrix(d, psi, pr, pc);
        else if (m.name == "compressed_Psi") read_matrix(d, cpsi, cr, cc);
        else if (m.name == "is_compressed") { if (d.need(1)) out.is_compressed = *d.p != 0; }
        else if (m.name == "maps") {                                  // std::vector<FFAT_Map<T,1>>: GetMapVal reads _shells.at(2)
            const uint64_t count = d.u64();
            for (uint64_t s = 0; s < count && d.ok; ++s) {
                Cur shell = d.sub(d.u64());
                if (s != 2) continue;
                have_shell = true;
                Chunk f;
                while (next_chunk(shell, f)) {
                    Cur& e = f.data;
                    if (f.name == "cellSize") out.cellsize = e.f64();
                    else if (f.name == "center") read_vec3(e, out.center1);
                    else if (f.name == "bboxLow") read_vec3(e, out.bboxlow);
                    else if (f.name == "bboxTop") read_vec3(e, out.bboxtop);
                    else if (f.name == "lowCorners") {
                        const uint64_t n = e.u64();
                        for (uint64_t i = 0; i < n && e.ok; ++i) { std::vector<double> v; if (read_vec3(e, v)) out.lowcorners.push_back(v); }
                    } else if (f.name == "N_elements") {
                        const uint64_t n = e.u64();
                        for (uint64_t i = 0; i < n && e.ok; ++i) { const int a = e.i32(), b = e.i32(); out.n_elements.push_back({a, b}); }
                    } else if (f.name == "strides") {
                        const uint64_t n = e.u64();
                        for (uint64_t i = 0; i < n && e.ok; ++i) out.strides.push_back(e.i32());
                    }
                    if (!e.ok) { err = "legacy .fatcube: truncated member \"" + f.name + "\" of shell 2"; return false; }
                }
                if (!shell.ok) { err = "legacy .fatcube: truncated shell"; return false; }
            }
        }
        if (!d.ok) { err = "legacy .fatcube: truncated member \"" + m.name + "\""; return false; }
    }
    if (!members.ok) { err = "legacy .fatcube: truncated member stream"; return false; }
    if (!have_shell) { err = "legacy .fatcube: fewer than 3 shells (GetMapVal reads _shells.at(2))"; return false; }
    // the protobuf form keeps ONE matrix: _compressed_Psi when _is_compressed, else _Psi (ffat_map_serialize.h:147-160)
    const std::vector<double>& src = out.is_compressed ? cpsi : psi;
    const int64_t rows = out.is_compressed ? cr : pr, cols = out.is_compressed ? cc : pc;
    for (int64_t c = 0; c < cols; ++c) out.psi.emplace_back(src.begin() + c * rows, src.begin() + (c + 1) * rows);
    return true;
}

void legacy_fatcube_encode(const FatcubeMap& m, std::string& out) {
    std::string shell;                                               // FFAT_Map<T,1>::InitSerialization (ffat_solver.h:440-452)
    { std::string d; lg_put_i32(d, m.modeid); lg_put_chunk(shell, "modeId", T_INT, d); }
    { std::string d; lg_put_f64(d, -1.0); lg_put_chunk(shell, "k", T_DBL, d); }                       // a shell's own _k stays at its default
    { std::string d; lg_put_f64(d, m.cellsize); lg_put_chunk(shell, "cellSize", T_DBL, d); }
    { std::string d; lg_put_u64(d, m.lowcorners.size()); for (auto& v : m.lowcorners) d += mat_bytes(v, 3, 1);
      lg_put_chunk(shell, "lowCorners", "St6vectorIN5Eigen6MatrixIdLi3ELi1ELi0ELi3ELi1EEESaIS2_EE", d); }
    { std::string d; lg_put_u64(d, m.n_elements.size()); for (auto& v : m.n_elements) { lg_put_i32(d, v[0]); lg_put_i32(d, v[1]); }
      lg_put_chunk(shell, "N_elements", "St6vectorISt4pairIiiESaIS1_EE", d); }
    { std::string d; lg_put_u64(d, m.strides.size()); for (int v : m.strides) lg_put_i32(d, v); lg_put_chunk(shell, "strides", "St6vectorIiSaIiEE", d); }
    int total = 0; for (auto& v : m.n_elements) total += v[0] * v[1];
    { std::string d; lg_put_i32(d, total); lg_put_chunk(shell, "N_elements_total", T_INT, d); }
    { std::string d; lg_put_u64(d, 0); lg_put_chunk(shell, "A", "St6vectorIN5Eigen6MatrixISt7complexIdELin1ELi1ELi0ELin1ELi1EEESaIS4_EE", d); }
    lg_put_chunk(shell, "center", T_V3, mat_bytes(m.center1, 3, 1));
    lg_put_chunk(shell, "bboxLow", T_V3, mat_bytes(m.bboxlow, 3, 1));
    lg_put_chunk(shell, "bboxTop", T_V3, mat_bytes(m.bboxtop, 3, 1));
    std::string body;                                                // FFAT_Map<T,3>::InitSerialization (:978-991)
    { std::string d; lg_put_i32(d, m.modeid); lg_put_chunk(body, "modeId", T_INT, d); }
    { std::string d; lg_put_f64(d, m.k); lg_put_chunk(body, "k", T_DBL, d); }
    { std::string d; lg_put_f64(d, m.cellsize); lg_put_chunk(body, "cellSize", T_DBL, d); }
    lg_put_chunk(body, "center", T_V3, mat_bytes(m.center3, 3, 1));
    { std::string d; lg_put_u64(d, 3); for (int s = 0; s < 3; ++s) { lg_put_u64(d, shell.size()); d += shell; }
      lg_put_chunk(body, "maps", "St6vectorIN14Gpu_Wavesolver8FFAT_MapIdLi1EEESaIS2_EE", d); }
    { std::string d; lg_put_u64(d, 3); for (int s = 0; s < 3; ++s) { lg_put_u64(d, m.n_elements.size()); for (auto& v : m.n_elements) { lg_put_i32(d, v[0]); lg_put_i32(d, v[1]); } }
      lg_put_chunk(body, "N_elements", "St6vectorIS_ISt4pairIiiESaIS1_EESaIS3_EE", d); }
    { std::string d; lg_put_u64(d, 3); lg_put_i32(d, 0); lg_put_i32(d, total); lg_put_i32(d, 2 * total); lg_put_chunk(body, "strides", "St6vectorIiSaIiEE", d); }   // shell offsets (:960-975)
    std::vector<double> flat; const size_t rows = m.psi.empty() ? 0 : m.psi[0].size();
    for (auto& c : m.psi) flat.insert(flat.end(), c.begin(), c.end());
    const std::string mat = mat_bytes(flat, (int64_t)rows, (int64_t)m.psi.size()), none = mat_bytes({}, 0, 0);
    lg_put_chunk(body, "Psi", T_MX, m.is_compressed ? none : mat);
    { std::string d; lg_put_i32(d, 3 * total); lg_put_chunk(body, "N_elements_total", T_INT, d); }
    { std::string d; lg_put_i32(d, total); lg_put_chunk(body, "N_directions", T_INT, d); }
    { std::string d; d.push_back(m.is_compressed ? 1 : 0); lg_put_chunk(body, "is_compressed", T_BOOL, d); }
    lg_put_chunk(body, "compressed_Psi", T_MX, m.is_compressed ? mat : none);
    std::string data; lg_put_u64(data, body.size()); data += body;
    out.clear();
    lg_put_chunk(out, "serial_map_ch3", "N14Gpu_Wavesolver8FFAT_MapIdLi3EEE", data);
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
