"""Policy-side consumer of the observation tensor: the network of the reference's
models/conv_to_fcnet_v2.py (ConvToFCNetv2) evaluated on uint8 observations that never leave the GPU.

    trunk   Conv2D(6, 3x3, stride 1, 'valid') -> ReLU -> flatten -> Dense(32) -> ReLU -> Dense(32) -> ReLU
            (conv_to_fcnet_v2.py:36-66, filters as train_baseline.py:104 configures them) -- ONE fused tcgen05 kernel
            behind the C ABI (`ssd_policy_features`, csrc/ssd_policy.cu) that reads the uint8 [B, N, 15, 15, 3] tensor
            `BatchedSSDEnv.step` wrote and folds the (x - 128) / 255 of map_env.py:199 into the first layer;
    head    LSTM(cell_size) -> logits / value (conv_to_fcnet_v2.py:68-92).  For the reference's cell_size 128 this is a second
            fused tcgen05 kernel (`ssd_policy_lstm_heads`: gate GEMM in tensor memory, cell update, head GEMM and -- in
            `act` -- Gumbel-max sampling of the action, one pass over the recurrent state).  Other cell sizes take the
            unfused route: library GEMMs (torch / cuBLAS, bf16 operands) around the one-pass cell update
            `ssd_policy_lstm_cell`.

The recurrent state (h, c) of the fused route lives in HBM in the kernel's tiled layout, shape [ceil(M/128), 8, 128, 16]
(group of 128 agents, block of 16 units, agent, unit): a warp of the cell update then moves 2 KB contiguous per access.  Treat
it as opaque between steps (as RLlib does with state_out -> state_in); `state_rows` / `state_from_rows` convert to and from
[M, cell_size].  Weights are numpy fp32 arrays in the Keras layouts (conv kernel [kh, kw, in, out], dense kernels [in, out], LSTM kernel
[in, 4u] / recurrent kernel [u, 4u] / bias [4u] with gates ordered i, f, c, o), so a checkpoint of the reference model
loads without transposition.  There is no CPU path: the trunk raises without the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

VIEW_RADIUS = 7
OBS_SIDE = 2 * VIEW_RADIUS + 1
CONV_FILTERS = 6
FEATURES = 32
FLAT = (OBS_SIDE - 2) * (OBS_SIDE - 2) * CONV_FILTERS   # 1014


def random_weights(num_outputs, cell_size=128, seed=0):
    """Random-init weights of the right shapes (normc for the dense layers as conv_to_fcnet_v2.py:63, glorot elsewhere)."""
    rng = np.random.RandomState(seed)

    def normc(shape, std=1.0):
        w = rng.randn(*shape).astype(np.float32)
        return (w * std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))).astype(np.float32)

    def glorot(shape, fan_in, fan_out):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    u = cell_size
    return {
        "conv_w": glorot((3, 3, 3, CONV_FILTERS), 27, 9 * CONV_FILTERS), "conv_b": (0.1 * rng.randn(CONV_FILTERS)).astype(np.float32),
        "fc1_w": normc((FLAT, FEATURES)), "fc1_b": (0.1 * rng.randn(FEATURES)).astype(np.float32),
        "fc2_w": normc((FEATURES, FEATURES)), "fc2_b": (0.1 * rng.randn(FEATURES)).astype(np.float32),
        "lstm_w": glorot((FEATURES, 4 * u), FEATURES, 4 * u), "lstm_u": glorot((u, 4 * u), u, 4 * u),
        "lstm_b": np.concatenate([np.zeros(u), np.ones(u), np.zeros(2 * u)]).astype(np.float32),   # unit forget bias
        "logits_w": glorot((u, num_outputs), u, num_outputs), "logits_b": np.zeros(num_outputs, np.float32),
        "value_w": glorot((u, 1), u, 1), "value_b": np.zeros(1, np.float32),
    }


class ConvToFCNet(object):
    """Forward pass of ConvToFCNetv2 for a batch of agents; `obs` is the uint8 tensor of the step path."""

    def __init__(self, weights, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SsdError("the policy trunk runs on a CUDA device only")
        if self.device.index is None:  # 'cuda' means the current device, not device 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        for name, shape in (("conv_w", (3, 3, 3, CONV_FILTERS)), ("conv_b", (CONV_FILTERS,)), ("fc1_w", (FLAT, FEATURES)),
                            ("fc1_b", (FEATURES,)), ("fc2_w", (FEATURES, FEATURES)), ("fc2_b", (FEATURES,))):
            if w[name].shape != shape:
                raise ValueError("%s has shape %s, expected %s" % (name, w[name].shape, shape))
        self.weights = w
        self._h = C.c_void_p()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib.ssd_policy_create(VIEW_RADIUS, self.device.index, ptr(w["conv_w"]), ptr(w["conv_b"]), ptr(w["fc1_w"]),
                                              ptr(w["fc1_b"]), ptr(w["fc2_w"]), ptr(w["fc2_b"]), C.byref(self._h)))
        self.cell_size = w["lstm_u"].shape[0] if "lstm_u" in w else 0
        self.num_outputs = w["logits_w"].shape[1] if "logits_w" in w else 0
        self._head = {}
        self._fused_head = False
        self._sample_seed, self._sample_counter = 0, 0
        if self.cell_size == 128 and 1 <= self.num_outputs <= 15:
            _lib.check(_lib.lib.ssd_policy_set_head(self._h, self.cell_size, self.num_outputs, ptr(w["lstm_w"]), ptr(w["lstm_u"]), ptr(w["lstm_b"]),
                                                    ptr(w["logits_w"]), ptr(w["logits_b"]), ptr(w["value_w"]), ptr(w["value_b"])))
            self._fused_head = True
        if self.cell_size:
            if self.cell_size % 8:
                raise ValueError("cell_size must be a multiple of 8")
            dev = lambda k, dt: torch.from_numpy(w[k]).to(self.device, dtype=dt).contiguous()
            self._head = {"lstm_w": dev("lstm_w", torch.bfloat16), "lstm_u": dev("lstm_u", torch.bfloat16), "lstm_b": dev("lstm_b", torch.float32),
                          "logits_w": dev("logits_w", torch.bfloat16), "logits_b": dev("logits_b", torch.float32),
                          "value_w": dev("value_w", torch.bfloat16), "value_b": dev("value_b", torch.float32)}

    def close(self):
        if self._h:
            _lib.lib.ssd_policy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def features(self, obs, out=None):
        """uint8 [..., 15, 15, 3] on the device -> float32 [M, 32] (M = product of the leading dims)."""
        if obs.dtype != torch.uint8 or obs.device != self.device or tuple(obs.shape[-3:]) != (OBS_SIDE, OBS_SIDE, 3):
            raise ValueError("obs must be a uint8 [..., %d, %d, 3] tensor on %s" % (OBS_SIDE, OBS_SIDE, self.device))
        obs = obs.contiguous()
        m = obs.numel() // (OBS_SIDE * OBS_SIDE * 3)
        if out is None:
            out = torch.empty((m, FEATURES), dtype=torch.float32, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib.ssd_policy_features(self._h, C.c_void_p(obs.data_ptr()), m, C.c_void_p(out.data_ptr()), stream))
        return out

    def initial_state(self, m):
        """Zero (h, c) for m agents, in the layout `forward` / `act` carry (tiled for the fused route, [m, cell_size] otherwise)."""
        shape = ((m + 127) // 128, self.cell_size // 16, 128, 16) if self._fused_head else (m, self.cell_size)
        z = torch.zeros(shape, dtype=torch.float32, device=self.device)
        return z, z.clone()

    @staticmethod
    def state_rows(state, m):
        """Tiled state [G, 8, 128, 16] -> [m, cell_size] (a copy)."""
        if state.dim() == 2:
            return state[:m]
        g, b, r, e = state.shape
        return state.permute(0, 2, 1, 3).reshape(g * r, b * e)[:m].contiguous()

    @staticmethod
    def state_from_rows(rows):
        """[m, cell_size] -> tiled state [ceil(m/128), cell_size/16, 128, 16], zero padded."""
        m, u = rows.shape
        g = (m + 127) // 128
        full = torch.zeros((g * 128, u), dtype=rows.dtype, device=rows.device)
        full[:m] = rows
        return full.reshape(g, 128, u // 16, 16).permute(0, 2, 1, 3).contiguous()

    def seed_sampling(self, seed):
        """Key of the Philox streams `act` samples from (counter = number of `act` calls since)."""
        self._sample_seed, self._sample_counter = int(seed) & (2 ** 64 - 1), 0

    def _fused(self, obs, h, c, sample):
        x = self.features(obs)
        m = x.shape[0]
        if h.dim() != 4 or h.shape[0] * 128 < m:
            raise ValueError("the fused route carries (h, c) in the tiled layout of initial_state / state_from_rows")
        h, c = h.contiguous(), c.contiguous()
        h_new, c_new = torch.empty_like(h), torch.empty_like(c)
        logits = torch.empty((m, self.num_outputs), dtype=torch.float32, device=self.device)
        value = torch.empty((m,), dtype=torch.float32, device=self.device)
        actions = torch.empty((m,), dtype=torch.int8, device=self.device) if sample else None
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)
        _lib.check(_lib.lib.ssd_policy_lstm_heads(self._h, p(x), p(h), p(c), p(h_new), p(c_new), p(logits), p(value), p(actions), m,
                                                  self._sample_seed, self._sample_counter & 0xFFFFFFFF, stream))
        if sample:
            self._sample_counter += 1
        return logits, value, h_new, c_new, actions

    def forward(self, obs, h, c):
        """-> (logits [M, A], value [M], h', c'), all float32; (h, c) in the layout of `initial_state`."""
        if not self._fused_head:
            return self.forward_unfused(obs, h, c)
        return self._fused(obs, h, c, sample=False)[:4]

    def forward_unfused(self, obs, h, c):
        """The same forward pass with library GEMMs: the trunk kernel, the gate GEMMs (cuBLAS, bf16 x bf16 -> fp32 accumulate), the
        one-pass cell update, the head GEMMs.  Any cell size that is a multiple of 8; (h, c) are [M, cell_size] here."""
        hd = self._head
        x16 = self.features(obs).to(torch.bfloat16)
        gates = torch.mm(x16, hd["lstm_w"]).addmm_(h.to(torch.bfloat16), hd["lstm_u"])   # bf16 [M, 4u], bias added in the cell kernel
        m, u = h.shape
        c_new, h_new = torch.empty_like(c), torch.empty_like(h)
        h16 = torch.empty((m, u), dtype=torch.bfloat16, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.lib.ssd_policy_lstm_cell(p(gates), p(hd["lstm_b"]), p(c.contiguous()), p(c_new), p(h_new), p(h16), m, u, stream))
        logits = torch.mm(h16, hd["logits_w"]).float() + hd["logits_b"]
        value = (torch.mm(h16, hd["value_w"]).float() + hd["value_b"]).squeeze(1)
        return logits, value, h_new, c_new

    def act(self, obs, h, c, generator=None):
        """Sample int8 actions [M] on the device for the next `BatchedSSDEnv.step`: inside the fused kernel (Philox streams,
        see `seed_sampling`), or with torch.multinomial (`generator`) on the unfused route."""
        if self._fused_head and generator is None:
            logits, value, h, c, a = self._fused(obs, h, c, sample=True)
            return a, value, h, c
        logits, value, h, c = self.forward(obs, h, c)
        a = torch.multinomial(torch.softmax(logits, dim=1), 1, generator=generator).squeeze(1).to(torch.int8)
        return a, value, h, c
