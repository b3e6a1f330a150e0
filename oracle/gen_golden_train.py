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
