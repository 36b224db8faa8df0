"""HWGATE hyper-parameters and skeleton graph: drop-in for the HWGATEParams class
of hwgat/models/model_params.py:243-403 (same constructor, same attributes, same
get_model_params() tuple).  The adjacency tensor is built on the GPU by kernel
K1a (hwgat_adjacency_build) instead of host numpy loops."""
from __future__ import annotations

import torch
import torch.nn as nn

from sl_hwgat_b200 import ops

# One window = 16 keypoints laid out by WindowCreate (dataTransform.py:426-455):
# slots 0-2 head, 3-5 arm, 6-15 hand.  The skeleton has a 3-star on the head,
# the arm chain into the wrist (slot 6), five 2-joint fingers fanning out of the
# wrist, and links between neighbouring fingers.  Same 25 undirected edges as
# model_params.py:261-369 (all four windows share them), written by structure
# rather than as a literal list.
def _window_edges():
    head = [(0, 1), (0, 2)]
    arm = [(0, 3), (3, 4), (4, 5), (5, 6)]
    thumb = [(6, 7)]
    finger_bases = [8, 10, 12, 14]
    palm = [(6, b) for b in finger_bases]
    finger = [(b, b + 1) for b in finger_bases]
    knuckles = [(finger_bases[i], finger_bases[i + 1]) for i in range(3)]
    tips = [(7, 9), (9, 11), (11, 13), (13, 15), (7, 15), (7, 11), (7, 13)]
    return [list(e) for e in head + arm + thumb + palm + finger + knuckles + tips]


class HWGATEParams():
    def __init__(self, dataset_params, input_dim, device=None) -> None:
        self.kp_dim = input_dim
        self.num_kps = 64
        self.temporal_dim = dataset_params['src_len']
        self.num_classes = dataset_params['num_class']
        self.embed_dim = 128
        self.temporal_patch_size = 2
        self.pe = True
        self.depths = [2, 2, 4]
        self.num_heads = [2, 4, 8]
        self.window_size = 16
        self.drop_rate = 0.1
        self.attn_drop_rate = 0.0
        self.ff_ratio = 2.
        self.norm_layer = nn.LayerNorm
        self.device = device
        self.edges = [_window_edges() for _ in range(self.num_kps // self.window_size)]
        # the reference hands the model a CPU float32 tensor (model_params.py:259)
        self.adj_mat = self.get_adj_mat().cpu()

    def set_window_size(self, window_size):
        """Larger keypoint windows (BASELINE configs[4]): the reference takes `window_size` at model_params.py:254
        and builds np.eye(W) + the 25 skeleton edges for the first K/W edge lists (model_params.py:373-400); this
        re-derives adj_mat the same way for W in {16, 32, 64}."""
        if self.num_kps % window_size != 0:
            raise ValueError("window size and number of kps are incompatible")
        self.window_size = window_size
        self.edges = [_window_edges() for _ in range(self.num_kps // window_size)]
        self.adj_mat = self.get_adj_mat().cpu()

    def _cuda_device(self):
        dev = torch.device(self.device) if self.device is not None else None
        if dev is None or dev.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("HWGATEParams builds the adjacency with a CUDA kernel and no CUDA "
                                   "device is available; there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    def get_adj_mat(self):
        """(nW, TP*W, TP*W) float32 (model_params.py:373-392), via kernel K1a."""
        return ops.adjacency_build(self.edges, self.window_size, self.temporal_patch_size, self._cuda_device())

    def get_adj(self, index):
        """(W, W) skeleton adjacency of window `index` (model_params.py:394-400):
        the same-frame block of the K1a output."""
        W = self.window_size
        return self.get_adj_mat()[index, :W, :W].cpu().numpy()

    def get_model_params(self):
        return (self.kp_dim, self.num_kps, self.temporal_dim, self.num_classes, self.embed_dim,
                self.temporal_patch_size, self.pe, self.depths, self.num_heads, self.window_size, self.adj_mat,
                self.drop_rate, self.attn_drop_rate, self.ff_ratio, self.norm_layer, self.device)


def _hand_edges(wrist):
    """12 edges of one 10-keypoint hand whose wrist is keypoint `wrist`: the thumb joint, four finger bases, a second
    joint per finger, and links between neighbouring finger bases."""
    bases = [wrist + 2 * i for i in range(1, 5)]
    return ([[wrist, wrist + 1]] + [[wrist, b] for b in bases] + [[b, b + 1] for b in bases] +
            [[bases[i], bases[i + 1]] for i in range(3)])


def _body_edges():
    """The 29-keypoint mediapipe skeleton of HGATE (same 34 undirected edges as model_params.py:422-457): 0-2 head,
    3-8 shoulders / elbows / wrists, 9-18 left hand, 19-28 right hand."""
    head = [[2, 0], [1, 0]]
    arms = [[0, 3], [0, 4], [3, 5], [4, 6], [5, 7], [6, 8]]
    return head + arms + [[7, 9]] + _hand_edges(9) + [[8, 19]] + _hand_edges(19)


class HGATEParams():
    """Drop-in for the HGATEParams class of hwgat/models/model_params.py:405-483 (same constructor, attributes and
    get_model_params() tuple); the (58, 58) block adjacency is built on the GPU by kernel K1a."""

    def __init__(self, dataset_params, input_dim, device=None) -> None:
        self.kp_dim = input_dim
        self.num_kps = 29
        self.temporal_dim = dataset_params['src_len']
        self.num_classes = dataset_params['num_class']
        self.embed_dim = 128
        self.temporal_patch_size = 2
        self.pe = True
        self.depths = [2, 2, 4]
        self.num_heads = [2, 4, 8]
        self.drop_rate = 0.1
        self.attn_drop_rate = 0.0
        self.ff_ratio = 2.
        self.norm_layer = nn.LayerNorm
        self.device = device
        self.edges = [_body_edges()]
        self.adj_mat = self.get_adj_mat().cpu()

    _cuda_device = HWGATEParams._cuda_device

    def get_adj_mat(self):
        """(TP*K, TP*K) float32 (model_params.py:461-474): one "window" of all K keypoints, via kernel K1a."""
        return ops.adjacency_build(self.edges, self.num_kps, self.temporal_patch_size, self._cuda_device())[0]

    def get_adj(self):
        """(K, K) skeleton adjacency (model_params.py:476-481)."""
        K = self.num_kps
        return self.get_adj_mat()[:K, :K].cpu().numpy()

    def get_model_params(self):
        return (self.kp_dim, self.num_kps, self.temporal_dim, self.num_classes, self.embed_dim,
                self.temporal_patch_size, self.pe, self.depths, self.num_heads, self.adj_mat, self.drop_rate,
                self.attn_drop_rate, self.ff_ratio, self.norm_layer, self.device)


def _frame_band(same, frames):
    """(frames*k, frames*k) float32: `same` (k, k) on the diagonal frame blocks, the identity between adjacent frames,
    zero further apart - the token order is frame * k + keypoint."""
    k = same.shape[0]
    eye_f = torch.eye(frames)
    step = torch.diag(torch.ones(frames - 1), 1) + torch.diag(torch.ones(frames - 1), -1) if frames > 1 \
        else torch.zeros(1, 1)
    return torch.kron(eye_f, same) + torch.kron(step, torch.eye(k))


class WGATEParams():
    """Drop-in for the WGATEParams class of hwgat/models/model_params.py:80-240 (same constructor, attributes and
    get_model_params() tuple).  adj_mat is (nW, F*W, F*W): the window's skeleton with self loops inside a frame, the
    identity between adjacent frames (model_params.py:204-229).  Host set-up work, built with torch on the CPU like
    the reference's numpy loops (it is F^2 blocks once per model, not on the step)."""

    def __init__(self, dataset_params, input_dim, device=None) -> None:
        self.kp_dim = input_dim
        self.num_kps = 64
        self.temporal_dim = dataset_params['src_len']
        self.num_classes = dataset_params['num_class']
        self.embed_dim = 128
        self.pe = True
        self.depths = 8
        self.num_heads = 8
        self.window_size = 16
        self.drop_rate = 0.1
        self.attn_drop_rate = 0.0
        self.ff_ratio = 2.
        self.norm_layer = nn.LayerNorm
        self.kp_norm = True
        self.device = device
        self.edges = [_window_edges() for _ in range(self.num_kps // self.window_size)]
        self.adj_mat = torch.as_tensor(self.get_adj_mat(), dtype=torch.float32)

    def get_adj_mat(self):
        return torch.stack([_frame_band(torch.as_tensor(self.get_adj(i), dtype=torch.float32), self.temporal_dim)
                            for i in range(len(self.edges))]).numpy()

    def get_adj(self, index):
        """(W, W) skeleton adjacency of window `index` with self loops (model_params.py:233-239)."""
        a = torch.eye(self.window_size)
        for i, j in self.edges[index]:
            a[i, j] = 1
            a[j, i] = 1
        return a.numpy()

    def get_model_params(self):
        return (self.kp_dim, self.num_kps, self.temporal_dim, self.num_classes, self.embed_dim, self.pe, self.depths,
                self.num_heads, self.window_size, self.ff_ratio, self.adj_mat, self.drop_rate, self.attn_drop_rate,
                self.norm_layer, self.device)


class GATEParams():
    """Drop-in for the GATEParams class of hwgat/models/model_params.py:5-77.  adj_mat is (F*29, F*29): the 34 body
    and hand edges inside every frame (NO self loops) and a link between the same keypoint of adjacent frames
    (model_params.py:59-74)."""

    def __init__(self, dataset_params, input_dim, device=None) -> None:
        self.kp_dim = input_dim
        self.num_kps = 29
        self.temporal_dim = dataset_params['src_len']
        self.num_classes = dataset_params['num_class']
        self.embed_dim = 128
        self.pe = True
        self.depths = 8
        self.num_heads = 8
        self.ff_ratio = 2.
        self.drop_rate = 0.1
        self.attn_drop_rate = 0.0
        self.norm_layer = nn.LayerNorm
        self.device = device
        self.edges = _body_edges()
        self.adj_mat = torch.as_tensor(self.get_adj(self.edges, self.temporal_dim, self.num_kps), dtype=torch.float32)

    def get_adj(self, spatial_links, num_fr, num_kp):
        same = torch.zeros(num_kp, num_kp)
        for i, j in spatial_links:
            same[i, j] = 1
            same[j, i] = 1
        return _frame_band(same, num_fr).numpy()

    def get_model_params(self):
        return (self.kp_dim, self.num_kps, self.temporal_dim, self.num_classes, self.embed_dim, self.pe, self.depths,
                self.num_heads, self.ff_ratio, self.adj_mat, self.drop_rate, self.attn_drop_rate, self.norm_layer,
                self.device)
