"""Development aid: compare the host-compiled eigensolver core with the oracle (expm)."""
import subprocess, sys, os, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import robchar_oracle as orc

def build_de(ctrl, normals, sigma, n):
    z = normals.reshape(normals.shape[:-1] + (n, 3)) * sigma
    d = ctrl[..., :n] + z[..., 0]
    e = np.hypot(1.0 + z[..., 1:, 1], z[..., 1:, 2])
    return d, e

def run(n, i, o, ctrl, normals, sigma, strided=0):
    d, e = build_de(ctrl, normals, sigma, n)
    T = np.abs(np.broadcast_to(ctrl[..., n], d.shape[:-1]))
    rec = np.concatenate([d, e, T[..., None]], axis=-1).reshape(-1, 2 * n)
    with open("/tmp/hs_in.bin", "wb") as f:
        f.write(struct.pack("4i", n, i, o, rec.shape[0])); f.write(rec.astype(np.float64).tobytes())
    subprocess.run(["/tmp/host_sim", "/tmp/hs_in.bin", "/tmp/hs_out.bin", str(strided)], check=True)
    return np.fromfile("/tmp/hs_out.bin").reshape(d.shape[:-1])

if __name__ == "__main__":
    g = np.load("tests/golden/replay_n7_0_6.npz")
    for name in ["replay_n4_0_2", "replay_n5_0_4", "replay_n6_0_3", "replay_n7_0_6"]:
        g = np.load(f"tests/golden/{name}.npz")
        n, i, o = map(int, g["nio"])
        ctrl = g["ctrl"][None, :, None, :]
        for strided in (0, 1, 2):
            f = run(n, i, o, ctrl, g["normals"], g["sigmas"][:, None, None, None, None], strided)
            ok = ~np.isnan(g["fids"])
            print(name, ["reg","strided","compact"][strided], "max |diff| vs reference run:", np.abs(f[ok] - g["fids"][ok]).max(),
                  "nan match", np.array_equal(np.isnan(f), np.isnan(g["fids"])))
    gl = np.load("tests/golden/replay_large_n.npz")
    for n in (10, 16, 32):
        for strided in (0, 1, 2):
            f = run(n, 0, n - 1, gl[f"n{n}_ctrl"][:, None, :], gl[f"n{n}_normals"], 0.05, strided)
            print("N", n, ["reg","strided","compact"][strided], np.abs(f - gl[f"n{n}_fids"]).max())
    # sweep statistics on synthetic controllers, N=7
    rs = np.random.RandomState(1)
    for n in (4, 7, 16):
        ctrl = orc.synthetic_controllers(300, n)
        normals = rs.standard_normal((300, 32, 3 * n))
        f = run(n, 0, n - 1, ctrl[:, None, :], normals, 0.05)
        fo = orc.fidelity_batch(ctrl[:, None, :], n, 0, n - 1, normals * 0.05)
        print("synthetic N", n, "max diff vs oracle", np.abs(f - fo).max())
