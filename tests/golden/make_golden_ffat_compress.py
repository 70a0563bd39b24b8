"""Generates tests/golden/ffat_compress.npz: FFAT_Map<T,3>::Compress (reference ffat_solver.h:1125-1178) carried out step
by step with the library the reference calls for it -- OpenCV (the Python build of it that this image ships; the reference
compiles that function only under USE_OPENCV and its C++ OpenCV is not here, so the C++ cannot be built in place):

    A_amp *= 255/maxAmp                       numpy, as Eigen does it: the scalar first, then one multiply per entry
    data.convertTo(data_s, CV_8U)             cv2.add(A, 0, dtype=CV_8U): the same saturate_cast<uchar>(double)
    cv::imwrite(jpg, quality) / cv::imread    cv2.imwrite / cv2.imread(IMREAD_GRAYSCALE), quality 65 (the default, :275-276)
    data_s.convertTo(data, CV_64F); A_amp *= maxAmp/255.

Kept per map: Psi, the bytes before and after the JPEG round trip, maxAmp per face, maxAmp_global and _compressed_Psi.  Run
HERE:    python tests/golden/make_golden_ffat_compress.py"""
import os
import sys
import tempfile
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openpbso_b200 import synth                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def compress(m, quality=65):
    psi = np.asarray(m["psi"], dtype=np.float64)
    pre = np.zeros(len(psi), np.uint8); post = np.zeros(len(psi), np.uint8); cpsi = np.zeros(len(psi))
    amp = np.empty(6); off = 0; gmax = -1.0
    with tempfile.TemporaryDirectory() as d:
        for fc, (nx, ny) in enumerate(np.asarray(m["n_elements"]).reshape(6, 2)):
            A = psi[off:off + nx * ny].reshape(nx, ny).copy()              # ConvertToImages (:1106-1122)
            mx = A.max(); gmax = max(gmax, mx); amp[fc] = mx
            A = A * (255 / mx)
            data_s = cv2.add(A, np.zeros_like(A), dtype=cv2.CV_8U)
            pre[off:off + nx * ny] = data_s.reshape(-1)
            name = os.path.join(d, "tmp-%u-%u-amp.jpg" % (m["modeid"], fc))
            cv2.imwrite(name, data_s, [cv2.IMWRITE_JPEG_QUALITY, quality])
            data_s = cv2.imread(name, cv2.IMREAD_GRAYSCALE)
            post[off:off + nx * ny] = data_s.reshape(-1)
            cpsi[off:off + nx * ny] = (data_s.astype(np.float64) * (mx / 255.)).reshape(-1)
            off += nx * ny
    return pre, post, amp, gmax, cpsi


def main():
    freqs = synth.mode_frequencies(3, 1004)
    maps = synth.ffat_maps(freqs, 2000, n=8)
    # one map with a face of negative values and one with an all-zero face: the cast's corner cases
    maps[1]["psi"] = np.asarray(maps[1]["psi"]).copy(); maps[1]["psi"][:64] *= -1.0
    maps[2]["psi"] = np.asarray(maps[2]["psi"]).copy(); maps[2]["psi"][64:128] = 0.0
    rows = [compress(m) for m in maps]
    np.savez_compressed(os.path.join(OUT, "ffat_compress.npz"), freqs=freqs, psi=np.stack([m["psi"] for m in maps]),
                        q8_pre=np.stack([r[0] for r in rows]), q8_post=np.stack([r[1] for r in rows]),
                        max_amp=np.stack([r[2] for r in rows]), max_amp_global=np.array([r[3] for r in rows]),
                        compressed_psi=np.stack([r[4] for r in rows]),
                        cast_in=np.array([0.5, 1.5, 2.5, 254.5, 255.5, 300, -3, np.nan, 1e12, -1e12, 3e9, 0.49999999999999994, 127.5, 128.5]),
                        cast_out=cv2.add(np.array([[0.5, 1.5, 2.5, 254.5, 255.5, 300, -3, np.nan, 1e12, -1e12, 3e9, 0.49999999999999994, 127.5, 128.5]]),
                                         np.zeros((1, 14)), dtype=cv2.CV_8U)[0])
    print("wrote ffat_compress.npz; bytes changed by JPEG:", [int((r[0] != r[1]).sum()) for r in rows], "of", len(rows[0][0]))


if __name__ == "__main__":
    main()
