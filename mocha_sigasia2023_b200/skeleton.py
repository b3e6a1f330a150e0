"""Skeleton topology constants of the 'mocha' layout and the graph matrices derived from them.

Restates (does not import) the reference's graph construction for the only configuration the hot
path uses (configs/config.yaml:42-43: layout 'mocha', strategy 'distance'):
  * joint tree           net/graph.py:65-79
  * body-part star       net/graph.py:207-218
  * hop distance         net/graph.py:290-301
  * column normalisation net/graph.py:304-312
  * 'distance' partition net/graph.py:127-131 / :250-254
  * pooling groups       net/graph.py:401-417, :459-461 ; un-pooling :543-559, :602-603
"""
from __future__ import annotations

import numpy as np

# parents of the 24 joints (Hips first); configs/dataset.yaml:63-64 / graph.py:68-75
JOINT_PARENTS = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 8, 9, 10, 11, 8, 13, 14, 8, 16, 17, 18, 0, 20, 21, 22]
JOINT_NAMES = [
    "Hips", "LeftUpLeg", "LeftLeg", "LeftFoot", "LeftToeBase", "Spine", "Spine1", "Spine2", "Spine3",
    "LeftShoulder", "LeftArm", "LeftForeArm", "LeftHand", "Neck", "Neck1", "Head", "RightShoulder",
    "RightArm", "RightForeArm", "RightHand", "RightUpLeg", "RightLeg", "RightFoot", "RightToeBase"]
# body parts: Spine, LeftLeg, LeftArm, Neck, RightArm, RightLeg
BODY_GROUPS = [[0, 5, 6, 7, 8], [1, 2, 3, 4], [9, 10, 11, 12], [13, 14, 15], [16, 17, 18, 19], [20, 21, 22, 23]]
# 25 bones = simulation root + joints (test_fullframework.py:101-102)
BONE_PARENTS = [-1] + [p + 1 for p in JOINT_PARENTS]
CONTACT_BONES = [5, 24]  # test_fullframework.py:104


def hop_distance(num_node: int, edges, max_hop: int) -> np.ndarray:
    adj = np.zeros((num_node, num_node))
    for i, j in edges:
        adj[i, j] = adj[j, i] = 1.0
    hop = np.full((num_node, num_node), np.inf)
    reach = [np.linalg.matrix_power(adj, d) > 0 for d in range(max_hop + 1)]
    for d in range(max_hop, -1, -1):
        hop[reach[d]] = d
    return hop


def distance_partition(num_node: int, edges, max_hop: int) -> np.ndarray:
    """A[k] holds the column-normalised adjacency restricted to node pairs at hop distance k."""
    hop = hop_distance(num_node, edges, max_hop)
    within = np.zeros((num_node, num_node))
    for d in range(max_hop + 1):
        within[hop == d] = 1.0
    col = within.sum(axis=0)
    scale = np.where(col > 0, 1.0 / np.where(col > 0, col, 1.0), 0.0)
    normed = within * scale[None, :]
    out = np.zeros((max_hop + 1, num_node, num_node))
    for d in range(max_hop + 1):
        out[d][hop == d] = normed[hop == d]
    return out


def joint_adjacency(max_hop: int = 2) -> np.ndarray:
    n = len(JOINT_PARENTS)
    edges = [(i, i) for i in range(n)] + [(i, JOINT_PARENTS[i]) for i in range(1, n)]
    return distance_partition(n, edges, max_hop).astype(np.float32)


def body_adjacency(max_hop: int = 1) -> np.ndarray:
    n = len(BODY_GROUPS)
    edges = [(i, i) for i in range(n)] + [(0, i) for i in range(1, n)]
    return distance_partition(n, edges, max_hop).astype(np.float32)


def pool_weight() -> np.ndarray:
    """[24, 6]; column p averages the joints of body part p."""
    w = np.zeros((len(JOINT_PARENTS), len(BODY_GROUPS)), dtype=np.float32)
    for p, group in enumerate(BODY_GROUPS):
        w[group, p] = 1.0
    return w / w.sum(axis=0, keepdims=True)


def unpool_weight() -> np.ndarray:
    """[6, 24]; copies a body-part feature to each of its joints (columns sum to 1)."""
    w = np.zeros((len(BODY_GROUPS), len(JOINT_PARENTS)), dtype=np.float32)
    for p, group in enumerate(BODY_GROUPS):
        w[p, group] = 1.0
    return w / w.sum(axis=0, keepdims=True)
