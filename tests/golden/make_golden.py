"""Generates the committed fixtures under tests/golden/ with the CPU oracle.  Run HERE (the build
container), where /root/reference/assets/ball.obj exists:

    python tests/golden/make_golden.py

Outputs
  cfg1_ball.npz      SURVEY 8(d) cfg1: ball.obj vertex 0 normal, 64 synthetic modes, FFAT transfer at
                     one listener, single PointForce, 173 buffers x 256 -> golden waveform
  fatcube/*.fatcube  small FFAT maps encoded by the stock python protobuf runtime (packed), incl.
                     modeId = 0 and zero-valued scalars (absent on the wire)
  fatcube_unpacked/  same content with UNPACKED repeated scalars (hand-encoded) -- parsers must accept both
  script_forces.npz  force state-machine script (point -> gaussian overlap -> clear -> sustained AR
                     start/update/end) and the oracle's buffers for it
"""
import os
import struct
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, fatcube          # noqa: E402
from openpbso_b200 import synth                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cfg1():
    V, F = orc.read_obj("/root/reference/assets/ball.obj")
    VN = orc.per_vertex_normals(V, F)
    vid = 0
    vn = VN[vid] / np.linalg.norm(VN[vid])         # tools/real_time_modal_sound.cpp:603 .normalized()
    M, K = 64, 3 * len(V)
    mat = synth.MATERIALS["low_damping"]
    freqs = synth.mode_frequencies(M, 1001)
    w2 = synth.omega_squared(freqs, mat["density"])
    U = synth.mode_shapes(M, K, 1001)
    space = orc.project_vertex(U, vid, vn)
    maps = synth.ffat_maps(freqs, 2000)
    listener = np.array([1e-3, 2e-3, 5.0])
    trans = orc.ffat_eval(maps, listener)[0]
    a, b = orc.build_ab(mat["density"], w2, mat["alpha"], mat["beta"])
    n_buf, BUF = 173, 256

    def render(scale):
        integ = orc.Integrator(synth.H, a, b)
        s = orc.Solver(integ, BUF)
        s.enqueue_trans(trans)
        s.enqueue_force(space * scale)
        return np.concatenate([s.step()[0] for _ in range(n_buf)])
    y0 = render(1.0)
    scale = 0.5e10 / np.max(np.abs(y0))            # peak |y| / 1e10 ~= 0.5
    y = render(scale)
    np.savez_compressed(os.path.join(OUT, "cfg1_ball.npz"), vn=vn, n_vertices=len(V), vid=vid, scale=scale,
                        listener=listener, space=space, trans=trans, y=y, u_vid=U[:, 3 * vid:3 * vid + 3])
    print("cfg1: peak/1e10 = %.4f, scale = %.4e" % (np.max(np.abs(y)) / 1e10, scale))


def small_maps():
    maps = []
    for mid, (n, R, c, k) in enumerate([(4, 1.0, (0, 0, 0), 1.25), (5, 2.0, (0.5, -0.25, 1.0), 0.0), (3, 0.75, (0, 0, 0), 7.5)]):
        g = synth.ffat_geometry(R, n, c)
        D = 6 * n * n
        psi = np.random.default_rng(77 + mid).uniform(0.1, 2.0, D)
        d = dict(g); d.update(k=k, psi=psi, modeid=mid, is_compressed=False)
        maps.append(d)
    return maps


def enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def ld(field, payload):
    return enc_varint((field << 3) | 2) + enc_varint(len(payload)) + payload


def unpacked_vec(v):
    return b"".join(enc_varint((1 << 3) | 1) + struct.pack("<d", float(x)) for x in v)


def unpacked_vec_i(v):
    return b"".join(enc_varint((1 << 3) | 0) + enc_varint(int(x)) for x in v)


def encode_unpacked(m):
    """Hand-rolled encoder with unpacked repeated scalars and an unknown field thrown in."""
    t1 = enc_varint((1 << 3) | 1) + struct.pack("<d", m["cellsize"])
    t1 += ld(2, b"".join(ld(1, unpacked_vec(r)) for r in np.asarray(m["lowcorners"])))
    t1 += ld(3, b"".join(ld(1, unpacked_vec_i(r)) for r in np.asarray(m["n_elements"])))
    t1 += ld(4, unpacked_vec_i(m["strides"]))
    t1 += ld(5, unpacked_vec(m["center1"])) + ld(6, unpacked_vec(m["bboxlow"])) + ld(7, unpacked_vec(m["bboxtop"]))
    t3 = b""
    if m["k"] != 0:
        t3 += enc_varint((1 << 3) | 1) + struct.pack("<d", m["k"])
    t3 += ld(2, unpacked_vec(m["center"])) + ld(3, t1)
    t3 += enc_varint((15 << 3) | 0) + enc_varint(12345)          # unknown field 15: must be skipped
    t3 += ld(5, ld(1, unpacked_vec(m["psi"])))
    if m["modeid"] != 0:
        t3 += enc_varint((6 << 3) | 0) + enc_varint(m["modeid"])
    return ld(1, t3)


def fatcubes():
    d1 = os.path.join(OUT, "fatcube"); d2 = os.path.join(OUT, "fatcube_unpacked")
    os.makedirs(d1, exist_ok=True); os.makedirs(d2, exist_ok=True)
    for m in small_maps():
        fatcube.save(os.path.join(d1, "mode-%d.fatcube" % m["modeid"]), m)
        with open(os.path.join(d2, "mode-%d.fatcube" % m["modeid"]), "wb") as f:
            f.write(encode_unpacked(m))
    # ListDirFiles filter (io.cpp:26-28): dot files and names without ".fatcube" are skipped
    open(os.path.join(d1, "freq_threshold.txt"), "w").write("15000\n")
    open(os.path.join(d1, ".hidden.fatcube"), "wb").write(b"\x00garbage")
    print("fatcube fixtures written")


def force_script():
    """One entry per step: (kind, arg).  Mirrors the message kinds of modal_solver.h:27-77."""
    N, BUF = 48, 256
    rng = np.random.default_rng(4242)
    mat = synth.MATERIALS["high_damping"]
    freqs = synth.mode_frequencies(N, 4243)
    a, b = synth.ab_from_material(freqs, mat)
    spaces = rng.standard_normal((8, N))
    trans = np.abs(rng.standard_normal((2, N))) + 0.1
    script = [("point", 0), ("none", 0), ("gauss", 1), ("point", 2), ("none", 0), ("trans", 0), ("none", 0),
              ("clear", 0), ("none", 0), ("point", 3), ("ar_start", 4), ("none", 0), ("arprm", 0), ("ar_data", 5),
              ("none", 0), ("ar_end", 6), ("none", 0), ("unit_transfer", 0), ("point", 7), ("use_transfer", 0),
              ("trans", 1), ("none", 0), ("none", 0)]
    integ = orc.Integrator(synth.H, a, b)
    s = orc.Solver(integ, BUF)
    ys = []; qns = []; produced = []
    for kind, arg in script:
        if kind == "point": s.enqueue_force(spaces[arg], orc.POINT)
        elif kind == "gauss": s.enqueue_force(spaces[arg], orc.GAUSSIAN, width_us=900.0)
        elif kind == "clear": s.enqueue_force(spaces[0], orc.POINT, flags=orc.F_CLEAR)
        elif kind == "ar_start": s.enqueue_force(spaces[arg], orc.AR, flags=orc.F_SUSTAIN_START)
        elif kind == "ar_data": s.enqueue_force(spaces[arg], orc.AR)
        elif kind == "ar_end": s.enqueue_force(spaces[arg], orc.AR, flags=orc.F_SUSTAIN_END)
        elif kind == "arprm": s.enqueue_arprm(0.7, 0.2, 0.002, 0.1)
        elif kind == "trans": s.enqueue_trans(trans[arg])
        elif kind == "unit_transfer": s.set_use_transfer(False)
        elif kind == "use_transfer": s.set_use_transfer(True)
        r = s.step()
        produced.append(r is not None)
        if r is not None:
            ys.append(r[0]); qns.append(r[1])
    np.savez_compressed(os.path.join(OUT, "script_forces.npz"), a=a, b=b, spaces=spaces, trans=trans,
                        kinds=np.array([k for k, _ in script]), args=np.array([x for _, x in script]),
                        produced=np.array(produced), y=np.array(ys), qnorm=np.array(qns))
    print("force script: %d steps, %d buffers" % (len(script), len(ys)))


if __name__ == "__main__":
    cfg1(); fatcubes(); force_script()
