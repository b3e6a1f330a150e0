"""Golden vectors for the training-side twins (tests/golden/train.npz): the reference's own `convert_YtilToX` and
`recon_criterion` (trainer.py:249-374) on seeded inputs. Build container only (imports /root/reference)."""
import os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stage_reference
REF = stage_reference.reference_root()
for p in ("", "etc", "motion", "preprocess", "net"):
    sys.path.insert(0, os.path.join(REF, p))
from mocha_sigasia2023_b200 import skeleton


def inputs(seed=3, B=2, T=8, V=24):
    rng = np.random.default_rng(seed)

    def win(J):
        y = rng.standard_normal((B, T, J, 15)).astype(np.float32)
        y[..., :3] *= 0.3
        # 6-D rotation channels: the first two columns of proper rotation matrices (what a trained network emits up to
        # noise). The reference's matrix FK keeps the first column un-normalised (txform.py:22-33) while the quaternion
        # kernels normalise, so the two only agree on orthonormal columns.
        q = rng.standard_normal((B, T, J, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        w, x, yq, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        c0 = np.stack([1 - 2 * (yq * yq + z * z), 2 * (x * yq + w * z), 2 * (x * z - w * yq)], -1)
        c1 = np.stack([2 * (x * yq - w * z), 1 - 2 * (x * x + z * z), 2 * (yq * z + w * x)], -1)
        y[..., 3:9] = np.stack([c0, c1], -1).reshape(B, T, J, 6).astype(np.float32)
        return y
    return win(V), win(V + 1)


if __name__ == "__main__":
    import trainer as ref_trainer
    Ytil, Ygt = inputs()
    parents = skeleton.BONE_PARENTS
    X = ref_trainer.convert_YtilToX(torch.from_numpy(Ytil), torch.from_numpy(Ygt[:, :, 0:1]), parents).numpy()
    loss = float(ref_trainer.recon_criterion(torch.from_numpy(Ytil), torch.from_numpy(Ygt), parents))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train.npz"), Ytil=Ytil, Ygt=Ygt, X=X, loss=np.float64(loss))
    print("train.npz", X.shape, loss)
