"""Independent `.fatcube` codec for the oracle -- TEST INFRASTRUCTURE ONLY.

Encodes/decodes ffat_map.proto (reference ffat_map.proto:12-51) with the stock `google.protobuf`
runtime through descriptors built in code (there is no protoc in the image).  It shares no code with
the product's hand-written wire codec (openpbso_b200/csrc/fatcube_codec.cpp), which is what makes the
loader parity test meaningful.  Field semantics follow ffat_map_serialize.h:90-254.
"""
import os
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ffat_map.proto"; fd.package = "ffat_map"; fd.syntax = "proto3"

    def msg(name, fields):
        m = fd.message_type.add(); m.name = name
        for (fname, num, ftype, label, tname) in fields:
            f = m.field.add(); f.name = fname; f.number = num; f.type = ftype; f.label = label
            if tname:
                f.type_name = ".ffat_map." + tname
    REP, OPT = _F.LABEL_REPEATED, _F.LABEL_OPTIONAL
    msg("vec", [("item", 1, _F.TYPE_DOUBLE, REP, None)])                      # :12-14
    msg("mat", [("item", 1, _F.TYPE_MESSAGE, REP, "vec")])                    # :17-19
    msg("vec_i", [("item", 1, _F.TYPE_INT32, REP, None)])                     # :21-23
    msg("mat_i", [("item", 1, _F.TYPE_MESSAGE, REP, "vec_i")])                # :26-28
    msg("ffat_map_t_1", [("cellsize", 1, _F.TYPE_DOUBLE, OPT, None),          # :30-38
                         ("lowcorners", 2, _F.TYPE_MESSAGE, OPT, "mat"),
                         ("n_elements", 3, _F.TYPE_MESSAGE, OPT, "mat_i"),
                         ("strides", 4, _F.TYPE_MESSAGE, OPT, "vec_i"),
                         ("center", 5, _F.TYPE_MESSAGE, OPT, "vec"),
                         ("bboxlow", 6, _F.TYPE_MESSAGE, OPT, "vec"),
                         ("bboxtop", 7, _F.TYPE_MESSAGE, OPT, "vec")])
    msg("ffat_map_t_3", [("k", 1, _F.TYPE_DOUBLE, OPT, None),                 # :40-47
                         ("center", 2, _F.TYPE_MESSAGE, OPT, "vec"),
                         ("shells", 3, _F.TYPE_MESSAGE, OPT, "ffat_map_t_1"),
                         ("is_compressed", 4, _F.TYPE_BOOL, OPT, None),
                         ("psi", 5, _F.TYPE_MESSAGE, OPT, "mat"),
                         ("modeid", 6, _F.TYPE_INT32, OPT, None)])
    msg("ffat_map_double", [("map", 1, _F.TYPE_MESSAGE, OPT, "ffat_map_t_3")])  # :49-51
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("ffat_map.ffat_map_double"))


_MSG = None


def _cls():
    global _MSG
    if _MSG is None:
        _MSG = _build()
    return _MSG


def encode(m):
    """dict -> bytes, mirroring FFAT_Map_Serialize_Double::Save (ffat_map_serialize.h:90-164).
    m: cellsize, lowcorners[6][3], n_elements[6][2], strides[6], center1[3], bboxlow[3], bboxtop[3],
       k, center[3], is_compressed, psi (1-D, column 0 of _Psi) or psi_cols (list of columns), modeid."""
    msg = _cls()()
    m3 = msg.map; m1 = m3.shells
    m1.cellsize = float(m["cellsize"])
    for row in np.asarray(m["lowcorners"], dtype=np.float64):
        m1.lowcorners.item.add().item.extend([float(x) for x in row])
    for row in np.asarray(m["n_elements"]):
        m1.n_elements.item.add().item.extend([int(x) for x in row])
    m1.strides.item.extend([int(x) for x in m["strides"]])
    m1.center.item.extend([float(x) for x in m["center1"]])
    m1.bboxlow.item.extend([float(x) for x in m["bboxlow"]])
    m1.bboxtop.item.extend([float(x) for x in m["bboxtop"]])
    m3.k = float(m["k"])
    m3.center.item.extend([float(x) for x in m["center"]])
    m3.is_compressed = bool(m.get("is_compressed", False))
    cols = m.get("psi_cols")
    if cols is None:
        cols = [m["psi"]]
    for col in cols:                      # column-major: one `vec` per column (:33-43)
        m3.psi.item.add().item.extend([float(x) for x in np.asarray(col, dtype=np.float64)])
    m3.modeid = int(m["modeid"])
    return msg.SerializeToString()


def decode(buf):
    """bytes -> dict, mirroring FFAT_Map_Serialize_Double::Load (ffat_map_serialize.h:166-254)."""
    msg = _cls()()
    msg.ParseFromString(buf)
    m3 = msg.map; m1 = m3.shells
    out = dict(
        cellsize=m1.cellsize,
        lowcorners=np.array([[v.item[j] for j in range(3)] for v in m1.lowcorners.item], dtype=np.float64),
        n_elements=np.array([[v.item[0], v.item[1]] for v in m1.n_elements.item], dtype=np.int32),
        strides=np.array(list(m1.strides.item), dtype=np.int32),
        center1=np.array(list(m1.center.item), dtype=np.float64),
        bboxlow=np.array(list(m1.bboxlow.item), dtype=np.float64),
        bboxtop=np.array(list(m1.bboxtop.item), dtype=np.float64),
        k=m3.k,
        center=np.array(list(m3.center.item), dtype=np.float64),
        is_compressed=m3.is_compressed,
        psi_cols=[np.array(list(v.item), dtype=np.float64) for v in m3.psi.item],
        modeid=m3.modeid,
    )
    out["psi"] = out["psi_cols"][0]       # GetMapVal reads column 0 only (ffat_solver.h:1203)
    return out


def save(path, m):
    with open(path, "wb") as f:
        f.write(encode(m))


def load(path):
    with open(path, "rb") as f:
        return decode(f.read())


def load_all(dirname):
    """FFAT_Map_Serialize_Double::LoadAll (ffat_map_serialize.h:267-279) + ListDirFiles
    (io.cpp:18-35): every non-dot entry whose full path contains '.fatcube', keyed by modeId."""
    out = {}
    for name in os.listdir(dirname):
        full = dirname + "/" + name
        if name[0] != "." and ".fatcube" in full and os.path.exists(full):
            m = load(full)
            out[m["modeid"]] = m
    return out


# ---------------------------------------------------------------------------------------------
# The LEGACY .fatcube form: libigl's igl::serialize of the FFAT_Map<T,3> object (ffat_solver.h:978-991 members,
# :1066-1085 Save / Load / LoadAll; the format is external/libigl/include/igl/serialize.h:452-480 chunk header,
# :700-1030 fundamental types / strings / std containers / Eigen matrices, :560-600 nested Serializable objects).
# [pinned: tests/golden/legacy_fatcube/*.fatcube are written by the reference's own FFAT_Map<double,3>::Save with libigl's own
#  serialize.h compiled in place (oracle/_ref, tests/golden/make_golden_legacy.py), and tests/test_oracle_vs_ref.py reads files
#  the reference has just written and compares with what the reference's own LoadAll + GetMapVal gives]
# ---------------------------------------------------------------------------------------------
import struct as _struct


class _Cur:
    def __init__(self, b, o=0, end=None):
        self.b = b; self.o = o; self.end = len(b) if end is None else end

    def take(self, n):
        if n < 0 or self.o + n > self.end:
            raise ValueError("legacy .fatcube: truncated")
        o = self.o; self.o += n
        return o

    def u64(self): return _struct.unpack_from("<Q", self.b, self.take(8))[0]
    def i32(self): return _struct.unpack_from("<i", self.b, self.take(4))[0]
    def f64(self): return _struct.unpack_from("<d", self.b, self.take(8))[0]
    def string(self): n = self.u64(); o = self.take(n); return self.b[o:o + n].decode("latin1")
    def sub(self, n): o = self.take(n); return _Cur(self.b, o, o + n)
    def more(self): return self.o < self.end

    def chunks(self):
        """[string name][string type][u64 size][data] until the end: yields (name, type, cursor over data)."""
        while self.more():
            name = self.string(); typ = self.string(); size = self.u64()
            yield name, typ, self.sub(size)

    def matrix(self):
        r = self.u64(); c = self.u64()                     # Eigen::Index = 8 bytes each, then column-major data
        o = self.take(8 * r * c)
        return np.frombuffer(self.b, dtype="<f8", count=r * c, offset=o).reshape(c, r).T.copy()


def is_legacy(buf):
    return len(buf) >= 22 and _struct.unpack_from("<Q", buf, 0)[0] == 14 and buf[8:22] == b"serial_map_ch3"


def decode_legacy(buf):
    """bytes -> the same dict decode() gives: shell 2's geometry (the one GetMapVal reads), k, centre, Psi."""
    body = None
    for name, _typ, cur in _Cur(buf).chunks():
        if name == "serial_map_ch3":
            body = cur                                      # igl::deserialize keeps the last match
    if body is None:
        raise ValueError("legacy .fatcube: no serial_map_ch3 object")
    top = {}; shell = None
    for name, _typ, d in body.sub(body.u64()).chunks():
        if name in ("modeId", "N_elements_total", "N_directions"): top[name] = d.i32()
        elif name in ("k", "cellSize"): top[name] = d.f64()
        elif name in ("center", "Psi", "compressed_Psi"): top[name] = d.matrix()
        elif name == "is_compressed": top[name] = d.b[d.take(1)] != 0
        elif name == "maps":
            shells = [d.sub(d.u64()) for _ in range(d.u64())]
            shell = {}
            for fname, _t, e in shells[2].chunks():         # GetMapVal reads _shells.at(2) (ffat_solver.h:1188-1203)
                if fname == "cellSize": shell[fname] = e.f64()
                elif fname in ("center", "bboxLow", "bboxTop"): shell[fname] = e.matrix()[:, 0]
                elif fname == "lowCorners": shell[fname] = np.array([e.matrix()[:, 0] for _ in range(e.u64())])
                elif fname == "N_elements": shell[fname] = np.array([[e.i32(), e.i32()] for _ in range(e.u64())], dtype=np.int32)
                elif fname == "strides": shell[fname] = np.array([e.i32() for _ in range(e.u64())], dtype=np.int32)
    comp = bool(top.get("is_compressed", False))
    P = top["compressed_Psi"] if comp else top["Psi"]
    out = dict(cellsize=shell["cellSize"], lowcorners=shell["lowCorners"], n_elements=shell["N_elements"], strides=shell["strides"],
               center1=shell["center"], bboxlow=shell["bboxLow"], bboxtop=shell["bboxTop"], k=top["k"], center=top["center"][:, 0],
               is_compressed=comp, psi_cols=[P[:, c].copy() for c in range(P.shape[1])], modeid=top["modeId"])
    out["psi"] = out["psi_cols"][0]
    return out


def load_any(path):
    """Either form of a .fatcube file."""
    with open(path, "rb") as f:
        buf = f.read()
    return decode_legacy(buf) if is_legacy(buf) else decode(buf)
