"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): float64 numpy restatement of the reference's policy network,
models/conv_to_fcnet_v2.py (ConvToFCNetv2), as configured by run_scripts/train_baseline.py:104 ("conv_filters": [[6, [3, 3], 1]],
cell_size 128) -- the checker of csrc/ssd_policy.cu / ssd_policy_head.cu.  TensorFlow is not installed in
this image, so the Keras layers are restated from their published definitions (tf.keras 2.0, the version the reference pins
in requirements.txt) and anchored on the reference's call sites:

  conv_to_fcnet_v2.py:46-55   Conv2D(6, 3x3, strides 1, padding "valid", activation = get_activation_fn(conv_activation) = ReLU);
                              kernel layout [kh, kw, in, out]: y[i, j, f] = sum_{di, dj, c} x[i+di, j+dj, c] * k[di, dj, c, f] + b[f]
  conv_to_fcnet_v2.py:56      flatten(): row-major over (row, col, filter) of the NHWC tensor
  conv_to_fcnet_v2.py:58-64   Dense(32, activation) twice: y = relu(x @ W + b), W [in, out]
  conv_to_fcnet_v2.py:75-79   tf.keras.layers.LSTM(cell_size): one step of LSTMCell.call with the tf.keras 2.0 defaults
                              activation = tanh, recurrent_activation = sigmoid, use_bias, unit_forget_bias only affects init:
                                  z = x @ kernel + h @ recurrent_kernel + bias,  (z_i, z_f, z_c, z_o) = split(z, 4)
                                  i = sigmoid(z_i); f = sigmoid(z_f); c' = f * c + i * tanh(z_c); o = sigmoid(z_o); h' = o * tanh(c')
  conv_to_fcnet_v2.py:82-88   logits = Dense(num_outputs, linear)(h'), values = Dense(1)(h')
  map_env.py:199              the network sees (rgb - 128.0) / 255.0

PARITY PIN: there is no TensorFlow here to generate golden vectors from, so this restatement is pinned by a hand-computed
vector (tests/test_host_logic.py::test_policy_ref_hand_vector) and by agreement with an independently written torch fp32
implementation (tests/test_policy_gpu.py); DESIGN.md says so."""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def features(w, obs_u8):
    """uint8 [M, 15, 15, 3] -> float64 [M, 32] (conv_to_fcnet_v2.py:36-66)."""
    x = (np.asarray(obs_u8).astype(np.float64) - 128.0) / 255.0
    win = np.lib.stride_tricks.sliding_window_view(x, (3, 3), axis=(1, 2))        # [M, 13, 13, C, kh, kw]
    y = np.einsum("mijcab,abcf->mijf", win, w["conv_w"].astype(np.float64)) + w["conv_b"].astype(np.float64)
    y = np.maximum(y, 0.0).reshape(x.shape[0], -1)
    y = np.maximum(y @ w["fc1_w"].astype(np.float64) + w["fc1_b"].astype(np.float64), 0.0)
    return np.maximum(y @ w["fc2_w"].astype(np.float64) + w["fc2_b"].astype(np.float64), 0.0)


def lstm_step(w, x, h, c):
    """One LSTMCell.call (see the module docstring).  x [M, 32], h / c [M, u] -> (h', c')."""
    z = x @ w["lstm_w"].astype(np.float64) + h @ w["lstm_u"].astype(np.float64) + w["lstm_b"].astype(np.float64)
    zi, zf, zc, zo = np.split(z, 4, axis=1)
    c2 = _sigmoid(zf) * c + _sigmoid(zi) * np.tanh(zc)
    h2 = _sigmoid(zo) * np.tanh(c2)
    return h2, c2


def forward(w, obs_u8, h, c):
    """(logits [M, A], values [M], h', c') of one step (conv_to_fcnet_v2.py:94-99)."""
    h2, c2 = lstm_step(w, features(w, obs_u8), np.asarray(h, np.float64), np.asarray(c, np.float64))
    logits = h2 @ w["logits_w"].astype(np.float64) + w["logits_b"].astype(np.float64)
    value = (h2 @ w["value_w"].astype(np.float64) + w["value_b"].astype(np.float64))[:, 0]
    return logits, value, h2, c2
