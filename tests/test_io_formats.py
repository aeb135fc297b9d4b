"""The reference's on-disk formats read / written by include/ptzcalib_io.hpp (data_io.cc:24-292): COLMAP text features and
matches, camera JSON.  CPU only."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_text_formats_and_camera_json(tmp_path):
    exe = str(tmp_path / "io_check")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "io_check.cpp"), "-o", exe], check=True)
    rng = np.random.default_rng(3)
    kp = rng.uniform(0, 1900, (5, 2)).astype(np.float32)
    desc = rng.uniform(0, 1, (5, 4)).astype(np.float32)
    feat = tmp_path / "a.jpg.txt"
    with open(feat, "w") as f:
        f.write("5 4\n")
        for i in range(5):
            f.write(f"{kp[i, 0]:.6f} {kp[i, 1]:.6f} 1.5 0.25 " + " ".join(f"{d:.6f}" for d in desc[i]) + "\n")
    matches = tmp_path / "matches.txt"
    with open(matches, "w") as f:
        f.write("a.jpg b.jpg\n0 3\n1 2\n4 4\n\nb.jpg c.png\n\nc.png a.jpg\n7 9\n\n")  # the middle block has no matches: dropped
    cams = {}
    for name in ("b", "a"):  # written out of order: the reference reads through std::map, i.e. sorted by key
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        R *= np.sign(np.linalg.det(R))
        cams[name] = dict(name=name, pos=[0, 0, 0], res=[1920, 1080], K=[1500.0, 0, 960, 0, 1510.0, 540, 0, 0, 1], R=R.ravel().tolist(),
                          t=[0.1, -0.2, 0.3], dist=[-0.1 if name == "b" else 0.0, 0, 0, 0, 0], distType="", version="2.0",
                          marker=dict(pix=[[0.25, 0.5], [0.75, 0.125]] if name == "a" else [], pos=[[1, 2, 3], [4, 5, 6]] if name == "a" else []))
    jin, jout = tmp_path / "cams.json", tmp_path / "out.json"
    with open(jin, "w") as f:
        json.dump(dict(cameras=cams), f, indent=4)
    r = subprocess.run([exe, str(feat), str(matches), str(jin), str(jout)], capture_output=True, text=True, check=True)
    lines = r.stdout.strip().split("\n")
    f0 = lines[0].split()
    assert f0[:3] == ["features", "5", "4"]
    got = np.array(f0[3:13], dtype=np.float32).reshape(5, 2)
    assert np.allclose(got, kp, atol=1e-4)
    assert lines[1].startswith("pairs 2 | a.jpg b.jpg 3 0:3 1:2 4:4 | c.png a.jpg 1 7:9")
    assert lines[2] == "json 1 2"
    assert lines[3] == "tracks 7 3 2 covis 0 1 2"
    assert lines[4] == "find 0 0 1 2 -1"
    assert lines[5] == "best c.png 1 |  0 | order a b"
    rt = lines[6].split()
    assert rt[1] == "1" and float(rt[2]) == 0.0 and rt[4] == "0"  # round trip exact; a missing camera makes ReadCamFromJson fail
    assert abs(float(rt[6]) - 480.0) < 1e-4 and abs(float(rt[7]) - 540.0) < 1e-4 and rt[9:] == ["1920", "1080"]  # marker pixels scaled by the resolution
    out = json.load(open(jout))  # what SaveToJson wrote is valid JSON with the reference's fields
    a = out["cameras"]["a"]
    assert list(a.keys()) == ["name", "pos", "res", "K", "R", "t", "dist", "distType", "marker", "version"]
    assert a["res"] == [1920, 1080] and a["K"] == cams["a"]["K"] and np.allclose(a["marker"]["pix"], cams["a"]["marker"]["pix"], atol=1e-6)
    Ra = np.array(cams["a"]["R"]).reshape(3, 3)
    assert np.allclose(a["pos"], -Ra.T @ np.array(cams["a"]["t"]))
    assert out["cameras"]["b"]["distType"] == "" and out["cameras"]["b"]["dist"][0] == -0.1  # 'k1' only for dist[0] >= 1e-5 (data_io.cc:143)
