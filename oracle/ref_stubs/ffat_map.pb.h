// TEST INFRASTRUCTURE ONLY -- hand-written stand-in for the protoc-generated ffat_map.pb.h (protoc and
// libprotobuf are absent in this image).  It offers the accessor subset that the reference's
// ffat_map_serialize.h:90-254 calls on the messages of ffat_map.proto:12-51 (add_item / item /
// item_size / mutable_x / x / set_x / SerializeToOstream / ParseFromIstream), over a small
// self-contained proto3 wire codec (canonical field order, packed repeated scalars on write, packed
// or unpacked accepted on read, zero scalars omitted, unknown fields skipped).  With it the
// reference's own Save/Load/LoadAll run unmodified; the wire format itself is pinned separately
// against the stock google.protobuf runtime (tests/test_host_logic.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <ostream>
#include <string>
#include <vector>

namespace ffat_map {
namespace wire {
inline void put_varint(std::string& s, uint64_t v) { while (v >= 0x80) { s.push_back((char)(v | 0x80)); v >>= 7; } s.push_back((char)v); }
inline void put_tag(std::string& s, int field, int wt) { put_varint(s, (uint64_t)field << 3 | wt); }
inline void put_f64(std::string& s, double d) { char b[8]; std::memcpy(b, &d, 8); s.append(b, 8); }
inline void put_len(std::string& s, int field, const std::string& body) { put_tag(s, field, 2); put_varint(s, body.size()); s += body; }
struct Reader {
    const uint8_t* p; const uint8_t* e; bool ok = true;
    Reader(const void* b, size_t n) : p((const uint8_t*)b), e((const uint8_t*)b + n) {}
    bool more() const { return ok && p < e; }
    uint64_t varint() { uint64_t v = 0; int sh = 0; while (p < e && sh < 64) { uint8_t c = *p++; v |= (uint64_t)(c & 0x7f) << sh; if (!(c & 0x80)) return v; sh += 7; } ok = false; return 0; }
    double f64() { if (e - p < 8) { ok = false; return 0; } double d; std::memcpy(&d, p, 8); p += 8; return d; }
    Reader sub() { uint64_t n = varint(); if (!ok || (uint64_t)(e - p) < n) { ok = false; return Reader(p, 0); } Reader r(p, (size_t)n); p += n; return r; }
    void skip(int wt) { if (wt == 0) varint(); else if (wt == 1) { if (e - p < 8) ok = false; else p += 8; } else if (wt == 2) sub(); else if (wt == 5) { if (e - p < 4) ok = false; else p += 4; } else ok = false; }
};
}  // namespace wire

class vec {
    std::vector<double> v_;
public:
    void add_item(double x) { v_.push_back(x); }
    double item(int i) const { return v_.at(i); }
    int item_size() const { return (int)v_.size(); }
    std::string bytes() const { std::string b; if (!v_.empty()) { std::string body; for (double x : v_) wire::put_f64(body, x); wire::put_len(b, 1, body); } return b; }
    bool parse(wire::Reader r) {
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7);
            if (f == 1 && wt == 2) { wire::Reader s = r.sub(); while (s.more()) v_.push_back(s.f64()); if (!s.ok) return false; }
            else if (f == 1 && wt == 1) v_.push_back(r.f64());
            else r.skip(wt); }
        return r.ok; }
};
class vec_i {
    std::vector<int32_t> v_;
public:
    void add_item(int32_t x) { v_.push_back(x); }
    int32_t item(int i) const { return v_.at(i); }
    int item_size() const { return (int)v_.size(); }
    std::string bytes() const { std::string b; if (!v_.empty()) { std::string body; for (int32_t x : v_) wire::put_varint(body, (uint64_t)(int64_t)x); wire::put_len(b, 1, body); } return b; }
    bool parse(wire::Reader r) {
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7);
            if (f == 1 && wt == 2) { wire::Reader s = r.sub(); while (s.more()) v_.push_back((int32_t)s.varint()); if (!s.ok) return false; }
            else if (f == 1 && wt == 0) v_.push_back((int32_t)r.varint());
            else r.skip(wt); }
        return r.ok; }
};
template <typename V> class mat_of {
    std::vector<V> v_;
public:
    V* add_item() { v_.emplace_back(); return &v_.back(); }
    const V& item(int i) const { return v_.at(i); }
    int item_size() const { return (int)v_.size(); }
    std::string bytes() const { std::string b; for (const V& x : v_) wire::put_len(b, 1, x.bytes()); return b; }
    bool parse(wire::Reader r) {
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7);
            if (f == 1 && wt == 2) { v_.emplace_back(); if (!v_.back().parse(r.sub())) return false; }
            else r.skip(wt); }
        return r.ok; }
};
typedef mat_of<vec> mat;
typedef mat_of<vec_i> mat_i;

// optional embedded message: proto3 emits it only when it was set (mutable_x() called or seen on the wire)
template <typename M> struct opt { M m; bool has = false; M* mut() { has = true; return &m; } };

class ffat_map_t_1 {
    double cellsize_ = 0; opt<mat> lowcorners_; opt<mat_i> n_elements_; opt<vec_i> strides_; opt<vec> center_, bboxlow_, bboxtop_;
public:
    void set_cellsize(double v) { cellsize_ = v; }
    double cellsize() const { return cellsize_; }
    mat* mutable_lowcorners() { return lowcorners_.mut(); }     const mat& lowcorners() const { return lowcorners_.m; }
    mat_i* mutable_n_elements() { return n_elements_.mut(); }   const mat_i& n_elements() const { return n_elements_.m; }
    vec_i* mutable_strides() { return strides_.mut(); }         const vec_i& strides() const { return strides_.m; }
    vec* mutable_center() { return center_.mut(); }             const vec& center() const { return center_.m; }
    vec* mutable_bboxlow() { return bboxlow_.mut(); }           const vec& bboxlow() const { return bboxlow_.m; }
    vec* mutable_bboxtop() { return bboxtop_.mut(); }           const vec& bboxtop() const { return bboxtop_.m; }
    std::string bytes() const {
        std::string b; uint64_t bits; std::memcpy(&bits, &cellsize_, 8);
        if (bits != 0) { wire::put_tag(b, 1, 1); wire::put_f64(b, cellsize_); }
        if (lowcorners_.has) wire::put_len(b, 2, lowcorners_.m.bytes());
        if (n_elements_.has) wire::put_len(b, 3, n_elements_.m.bytes());
        if (strides_.has) wire::put_len(b, 4, strides_.m.bytes());
        if (center_.has) wire::put_len(b, 5, center_.m.bytes());
        if (bboxlow_.has) wire::put_len(b, 6, bboxlow_.m.bytes());
        if (bboxtop_.has) wire::put_len(b, 7, bboxtop_.m.bytes());
        return b; }
    bool parse(wire::Reader r) {
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7); bool ok = true;
            if (f == 1 && wt == 1) cellsize_ = r.f64();
            else if (f == 2 && wt == 2) ok = lowcorners_.mut()->parse(r.sub());
            else if (f == 3 && wt == 2) ok = n_elements_.mut()->parse(r.sub());
            else if (f == 4 && wt == 2) ok = strides_.mut()->parse(r.sub());
            else if (f == 5 && wt == 2) ok = center_.mut()->parse(r.sub());
            else if (f == 6 && wt == 2) ok = bboxlow_.mut()->parse(r.sub());
            else if (f == 7 && wt == 2) ok = bboxtop_.mut()->parse(r.sub());
            else r.skip(wt);
            if (!ok) return false; }
        return r.ok; }
};

class ffat_map_t_3 {
    double k_ = 0; opt<vec> center_; opt<ffat_map_t_1> shells_; bool is_compressed_ = false; opt<mat> psi_; int32_t modeid_ = 0;
public:
    void set_k(double v) { k_ = v; }                       double k() const { return k_; }
    vec* mutable_center() { return center_.mut(); }        const vec& center() const { return center_.m; }
    ffat_map_t_1* mutable_shells() { return shells_.mut(); } const ffat_map_t_1& shells() const { return shells_.m; }
    void set_is_compressed(bool v) { is_compressed_ = v; } bool is_compressed() const { return is_compressed_; }
    mat* mutable_psi() { return psi_.mut(); }              const mat& psi() const { return psi_.m; }
    void set_modeid(int32_t v) { modeid_ = v; }            int32_t modeid() const { return modeid_; }
    std::string bytes() const {
        std::string b; uint64_t bits; std::memcpy(&bits, &k_, 8);
        if (bits != 0) { wire::put_tag(b, 1, 1); wire::put_f64(b, k_); }
        if (center_.has) wire::put_len(b, 2, center_.m.bytes());
        if (shells_.has) wire::put_len(b, 3, shells_.m.bytes());
        if (is_compressed_) { wire::put_tag(b, 4, 0); wire::put_varint(b, 1); }
        if (psi_.has) wire::put_len(b, 5, psi_.m.bytes());
        if (modeid_ != 0) { wire::put_tag(b, 6, 0); wire::put_varint(b, (uint64_t)(int64_t)modeid_); }
        return b; }
    bool parse(wire::Reader r) {
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7); bool ok = true;
            if (f == 1 && wt == 1) k_ = r.f64();
            else if (f == 2 && wt == 2) ok = center_.mut()->parse(r.sub());
            else if (f == 3 && wt == 2) ok = shells_.mut()->parse(r.sub());
            else if (f == 4 && wt == 0) is_compressed_ = r.varint() != 0;
            else if (f == 5 && wt == 2) ok = psi_.mut()->parse(r.sub());
            else if (f == 6 && wt == 0) modeid_ = (int32_t)r.varint();
            else r.skip(wt);
            if (!ok) return false; }
        return r.ok; }
};

class ffat_map_double {
    opt<ffat_map_t_3> map_;
public:
    ffat_map_t_3* mutable_map() { return map_.mut(); }
    const ffat_map_t_3& map() const { return map_.m; }
    bool SerializeToOstream(std::ostream* os) const {
        std::string b; if (map_.has) wire::put_len(b, 1, map_.m.bytes());
        os->write(b.data(), (std::streamsize)b.size()); return (bool)*os; }
    bool ParseFromIstream(std::istream* is) {
        std::string b((std::istreambuf_iterator<char>(*is)), std::istreambuf_iterator<char>());
        map_ = opt<ffat_map_t_3>();
        wire::Reader r(b.data(), b.size());
        while (r.more()) { uint64_t t = r.varint(); int f = (int)(t >> 3), wt = (int)(t & 7);
            if (f == 1 && wt == 2) { if (!map_.mut()->parse(r.sub())) return false; }
            else r.skip(wt); }
        return r.ok; }
};
}  // namespace ffat_map
