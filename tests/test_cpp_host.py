"""The C++ host mirror (include/surf_b200.hpp) compiles against the C ABI with plain g++ and, on a
GPU, produces the oracle's results when driven like the reference's surfcmd (Logic/surfcmd.cxx)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "surfcmd_b200.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "surfcmd_b200")


def _build(src=SRC, exe=EXE):
    import surffeatures_b200 as sb

    sb.load_library()
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    libdir = os.path.dirname(sb.LIB_PATH)
    cmd = [gxx, "-std=c++11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", libdir, "-lsurf_b200", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64",
           "-L/usr/local/cuda/lib64", "-lcudart"]
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return exe


def test_cpp_host_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = _build()
    raw = tmp_path / "f.raw"
    np.zeros((2, 48, 64), np.uint8).tofile(raw)
    r = subprocess.run([exe, str(raw), "64", "48", "2"], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "no CPU path" in r.stderr  # SURF_ERR_NO_DEVICE, not a silent fallback
    else:
        assert r.returncode == 0


@pytest.mark.gpu
def test_cpp_host_matches_oracle(tmp_path, oracle):
    from surffeatures_b200 import speckle

    exe = _build()
    frames = speckle.speckle_stream(3, 240, 320, seed=4, fresh=0.0)
    raw = tmp_path / "f.raw"
    frames.tofile(raw)
    r = subprocess.run([exe, str(raw), "320", "240", "3", "400"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 3
    prev = None
    for i, line in enumerate(lines):
        tok = line.split()
        kd = oracle.detect(frames[i], oracle.params(400, 4, 2))
        kc, dc = oracle.compute(frames[i], oracle.params(100, 4, 2), kd)
        assert int(tok[3]) == len(kc)
        assert abs(float(tok[5]) - float(dc.astype(np.float64).sum())) < 1e-2
        if prev is not None:
            m = oracle.knn_match(dc, prev, 2)
            good = int((m["distance"][:, 0] < np.float32(0.8) * m["distance"][:, 1]).sum())
            assert abs(int(tok[7]) - good) <= 1
        prev = dc


SRC_MHA = os.path.join(ROOT, "tests", "cpp", "surfcmd_mha_b200.cpp")
EXE_MHA = os.path.join(ROOT, "tests", "cpp", "build", "surfcmd_mha_b200")


def _write_case(tmp_path):
    from surffeatures_b200 import mha, speckle

    stream = speckle.speckle_stream(9, 480, 640, seed=31, fresh=0.0)
    sets = {"train": stream[:6], "query": stream[[2, 8]], "bogus": speckle.speckle_batch(2, 480, 640, 800)}
    files = {}
    for name, fr in sets.items():
        files[name] = str(tmp_path / f"{name}.mha")
        mha.write_mha(files[name], fr)
    yy, xx = np.mgrid[0:480, 0:640]
    mask = ((xx - 320) ** 2 + (yy + 40) ** 2 < 470 ** 2).astype(np.uint8) * 255
    mpath = str(tmp_path / "mask.raw")
    mask.tofile(mpath)
    return files, sets, mask, mpath


def test_cpp_logic_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    """The reference's computeAll() body compiles unchanged against surfb200::SurfFeaturesLogic."""
    import torch

    exe = _build(SRC_MHA, EXE_MHA)
    files, _, _, mpath = _write_case(tmp_path)
    r = subprocess.run([exe, files["bogus"], files["train"], files["query"], mpath, "400"], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "no CPU path" in r.stderr
    else:
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_surfcmd_flow_matches_oracle(tmp_path, oracle):
    from oracle import mha_oracle as mo
    from surffeatures_b200 import mha

    exe = _build(SRC_MHA, EXE_MHA)
    files, sets, mask, mpath = _write_case(tmp_path)
    r = subprocess.run([exe, files["bogus"], files["train"], files["query"], mpath, "400"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    ratios = mha.REFERENCE_CROP_RATIOS
    cmask = mo.crop(mask, ratios)
    cx, cy = oracle.mask_centroid(cmask)
    P = oracle.params(400, 4, 2, True, False)
    feats = {n: [oracle.detect_and_compute(mo.crop(f, ratios), P, mask=cmask) for f in mo.read_mha(files[n])["images"]]
             for n in sets}
    head = lines[0].split()
    assert (int(head[1]), int(head[2])) == (cx, cy)
    assert int(head[4]) == sum(len(k) for k, _ in feats["train"])
    assert len(lines) == 2 + len(feats["query"])
    # last line: SurfFeaturesLogic::ransac on a plane through query frame 0's keypoints (row f4)
    from oracle import pose_oracle as po

    qk0 = feats["query"][0][0][:200]
    pts = np.column_stack([qk0["x"].astype(np.float64), qk0["y"].astype(np.float64),
                           0.05 * qk0["x"].astype(np.float64) - 0.02 * qk0["y"].astype(np.float64)
                           + np.where(np.arange(len(qk0)) % 5 == 0, 30.0, 0.0)])
    rc, oinl, oplane, _, _ = po.ransac_plane(po.Rand(1), pts, margin=1.0, iterations=1000)
    tok = lines[-1].split()
    assert tok[0] == "ransac" and int(tok[2]) == rc == 0 and int(tok[4]) == len(qk0)
    assert int(tok[6]) == len(oinl) and [int(t) for t in tok[8:11]] == oplane.tolist()
    lines = lines[:-1]
    for i, (qk, qd) in enumerate(feats["query"]):
        want, res = oracle.correspond(qk, qd, [k for k, _ in feats["train"]], [d for _, d in feats["train"]],
                                      [d for _, d in feats["bogus"]], cx, cy)
        tok = lines[1 + i].split()
        got = {tok[j]: int(tok[j + 1]) for j in range(0, len(tok), 2)}
        assert got["keypoints"] == len(qk)
        assert (got["filtered"], got["hough"], got["best"], got["count"]) == tuple(int(v) for v in res)
        assert got["alive"] == int((want["trainIdx"] >= 0).sum())
        assert got["with_best"] == int((want["imgIdx"] == res[2]).sum())
    assert lines[1].split()[11] == "2"          # query frame 0 is train frame 2
