"""Writes tests/golden/golden.json (+ golden_samples.npz): outputs of the CPU oracle (oracle/mm_oracle.cpp) for the
cases of tests/cases.py, after checking that the independent numpy transcription (oracle/np_oracle.py) agrees bit for
bit on the NP_CASES, and — for the REF_SHADER_CASES — the image produced by the reference's OWN shader source compiled
here (oracle/ref_shader.py -> oracle/_ref/libref_shader.so, needs /root/reference), stored as `ref_shader_image`.
These pin the oracle against regressions and against the reference shader on machines that have neither the reference
tree nor the built _ref library, and give the GPU tests a fixture that does not need the oracle's .so.
Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import mirror_maze_b200 as mm          # noqa: E402  (host surface only; no GPU work)
from oracle import np_oracle, oracle, ref_shader   # noqa: E402
from cases import CASES, NP_CASES, REF_SHADER_CASES, build_case   # noqa: E402

STRIDE = 61


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


if __name__ == "__main__":
    noise = mm.load_noise()
    out, samples = {}, {}
    for name in CASES:
        sc, u, p, chunks = build_case(mm, name)
        img, cnt, dbg = oracle.render(sc, noise, u, p, chunks, debug=True)
        if name in NP_CASES:
            img2, cnt2, dbg2 = np_oracle.render(sc, noise, u, p, chunks)
            for k in dbg:
                assert dbg[k].tobytes() == dbg2[k].tobytes(), (name, k)
            assert img.tobytes() == img2.tobytes(), name
            for k in cnt2:
                assert cnt[k] == cnt2[k], (name, k)
        cnt.pop("literal_rays")
        out[name] = {"counters": cnt, "image": digest(img), **{k: digest(v) for k, v in dbg.items()},
                     "np_checked": name in NP_CASES, "planes": int(sc.n_planes), "nodes": int(sc.n_nodes)}
        if name in REF_SHADER_CASES:
            assert ref_shader.available(), "build oracle/_ref first (make -C oracle, needs /root/reference)"
            ref_img = ref_shader.render(sc, noise, u, p, chunks)
            out[name]["ref_shader_image"] = digest(ref_img)
            assert ref_img.tobytes() == img.tobytes(), f"{name}: oracle differs from the reference's own shader"
        for k, v in dbg.items():
            samples[f"{name}.{k}"] = v[::STRIDE].copy()
        samples[f"{name}.image_rows"] = img[:: max(1, img.shape[0] // 8)].copy()
        print(name, cnt)
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_samples.npz"), **samples)
