"""Oracle: CRNN recognizer (conv stack + STN + BiLSTM + greedy CTC), fp32 on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinning: the graph wiring, the STN sampler and the
CTC padding are checked against the reference's OWN source (``build_model``, ``_transform``,
``CTCDecoder``; recognition.py:54-350) executed on ``oracle/keras_shim.py`` by
``oracle/validate_against_reference.py`` (softmax equal to 4e-7, labels identical; fixture
tests/golden/crnn.npz).  PARITY UNPINNED against TensorFlow's kernels (not installable here): the
arithmetic inside each layer follows the Keras semantics listed in SURVEY.md Appendix B (LSTM gate
order i,f,c,o with sigmoid/tanh; ``go_backwards`` outputs kept in processing order;
BatchNormalization eps=1e-3; greedy CTC with repeat merging, blank = last class, -1 padding).

Weights: flat dict with Keras-style names and Keras layouts
  conv_N.kernel (kh,kw,cin,cout), conv_N.bias, bn_N.{gamma,beta,moving_mean,moving_variance},
  stn.conv_a / stn.conv_b / stn.dense_a / stn.dense_b (.kernel/.bias),
  fc_9, lstm_10, lstm_10_back, lstm_11, lstm_11_back (.kernel/.recurrent_kernel/.bias), fc_12.
"""
import numpy as np
import torch
import torch.nn.functional as F

KERAS_BN_EPS = 1e-3           # keras.layers.BatchNormalization default (recognition.py:226,234,242)
STEPS_TO_DISCARD = 2          # DEFAULT_BUILD_PARAMS["rnn_steps_to_discard"], recognition.py:20
ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"   # recognition.py:25
BLANK = len(ALPHABET)         # recognition.py:376


def _t(weights):
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in weights.items()}


def _conv_relu(w, x, name, k):
    kern = w[name + ".kernel"].permute(3, 2, 0, 1)      # HWIO -> OIHW
    return F.relu(F.conv2d(x, kern, w[name + ".bias"], padding=k // 2))


def _bn(w, x, name):
    return F.batch_norm(x, w[name + ".moving_mean"], w[name + ".moving_variance"],
                        w[name + ".gamma"], w[name + ".beta"], training=False, eps=KERAS_BN_EPS)


def stn_theta(w, feat):
    """Localisation net, recognition.py:268-278.  feat: (B,512,50,7) -> theta (B,6)."""
    y = _conv_relu(w, feat, "stn.conv_a", 5)
    y = _conv_relu(w, y, "stn.conv_b", 5)
    y = y.permute(0, 2, 3, 1).reshape(y.shape[0], -1)   # Flatten of a channels-last tensor
    y = F.relu(y @ w["stn.dense_a.kernel"] + w["stn.dense_a.bias"])
    return y @ w["stn.dense_b.kernel"] + w["stn.dense_b.bias"]


def stn_sample(feat_nhwc, theta):
    """``_transform``, recognition.py:73-166.  feat_nhwc: (B,Hh,Ww,C); theta: (B,6).

    Quirks kept on purpose: coordinates are scaled by W and H (not W-1/H-1, lines
    109-110) and the bilinear weights use the *clipped* corner indices (112-124,
    144-152), so samples past the last row/column get zero total weight.
    """
    B, Hh, Ww, C = feat_nhwc.shape
    dev = feat_nhwc.device
    xs = torch.linspace(-1.0, 1.0, Ww, device=dev)
    ys = torch.linspace(-1.0, 1.0, Hh, device=dev)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    grid = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.ones(Hh * Ww, device=dev)], 0)   # (3, Hh*Ww)
    tg = theta.reshape(B, 2, 3) @ grid                                              # (B,2,P)
    x = 0.5 * (tg[:, 0] + 1.0) * float(Ww)
    y = 0.5 * (tg[:, 1] + 1.0) * float(Hh)
    x0 = torch.floor(x).to(torch.int64)
    y0 = torch.floor(y).to(torch.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    x0 = x0.clamp(0, Ww - 1); x1 = x1.clamp(0, Ww - 1)
    y0 = y0.clamp(0, Hh - 1); y1 = y1.clamp(0, Hh - 1)
    flat = feat_nhwc.reshape(B, Hh * Ww, C)

    def gather(yy, xx):
        idx = (yy * Ww + xx).unsqueeze(-1).expand(-1, -1, C)
        return torch.gather(flat, 1, idx)

    x0f, x1f, y0f, y1f = x0.float(), x1.float(), y0.float(), y1.float()
    wa = ((x1f - x) * (y1f - y)).unsqueeze(-1)
    wb = ((x1f - x) * (y - y0f)).unsqueeze(-1)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(-1)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(-1)
    out = wa * gather(y0, x0) + wb * gather(y1, x0) + wc * gather(y0, x1) + wd * gather(y1, x1)
    return out.reshape(B, Hh, Ww, C)


def lstm(w, x, name, go_backwards=False):
    """keras.layers.LSTM(return_sequences=True) with TF2 defaults; x: (B,T,F) -> (B,T,U)."""
    W, U, b = w[name + ".kernel"], w[name + ".recurrent_kernel"], w[name + ".bias"]
    units = U.shape[0]
    if go_backwards:
        x = torch.flip(x, [1])
    B, T, _ = x.shape
    h = torch.zeros(B, units)
    c = torch.zeros(B, units)
    xz = x @ W + b
    outs = []
    for t in range(T):
        z = xz[:, t] + h @ U
        zi, zf, zc, zo = torch.split(z, units, dim=1)
        c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zc)
        h = torch.sigmoid(zo) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)      # processing order (NOT re-reversed), recognition.py:298-319


def crnn_features(weights, crops):
    """crops: (B,31,200), (B,31,200,1) or -- color model -- (B,31,200,3) float32 in [0,1].  Returns bn_7's output (B,512,50,7)."""
    w = _t(weights)
    x = torch.as_tensor(np.asarray(crops), dtype=torch.float32)
    if x.dim() == 3:
        x = x[..., None]                                  # gray crops (B,31,200) -> one channel
    # Permute((2,1,3)) then reverse axis 2 (recognition.py:215-216): (B,200,31,C), x[b,t,j] = crop[b,30-j,t];
    # C = 1, or 3 for build_model(color=True) (recognition.py:214)
    x = torch.flip(x.permute(0, 2, 1, 3), [2]).permute(0, 3, 1, 2).contiguous()   # NCHW with H=200, W=31
    x = _conv_relu(w, x, "conv_1", 3)
    x = _conv_relu(w, x, "conv_2", 3)
    x = _conv_relu(w, x, "conv_3", 3)
    x = F.max_pool2d(_bn(w, x, "bn_3"), 2, 2)
    x = _conv_relu(w, x, "conv_4", 3)
    x = _conv_relu(w, x, "conv_5", 3)
    x = F.max_pool2d(_bn(w, x, "bn_5"), 2, 2)
    x = _conv_relu(w, x, "conv_6", 3)
    x = _conv_relu(w, x, "conv_7", 3)
    x = _bn(w, x, "bn_7")
    return x


def crnn_logits(weights, crops, return_intermediates=False):
    """Returns the softmax outputs after discarding 2 steps: (B,48,37) float32."""
    w = _t(weights)
    feat = crnn_features(weights, crops)                 # (B,512,50,7)
    if "stn.conv_a.kernel" in w:
        theta = stn_theta(w, feat)
        warped = stn_sample(feat.permute(0, 2, 3, 1).contiguous(), theta)   # (B,50,7,512)
    else:                                                # build_model(stn=False), recognition.py:243: no transformer
        theta, warped = None, feat.permute(0, 2, 3, 1).contiguous()
    B = warped.shape[0]
    seq = warped.reshape(B, warped.shape[1], -1)         # Reshape -> (B,50,3584), feature = h*512+c
    seq = F.relu(seq @ w["fc_9.kernel"] + w["fc_9.bias"])
    l1 = lstm(w, seq, "lstm_10") + lstm(w, seq, "lstm_10_back", go_backwards=True)
    l2 = torch.cat([lstm(w, l1, "lstm_11"), lstm(w, l1, "lstm_11_back", go_backwards=True)], -1)
    logits = l2 @ w["fc_12.kernel"] + w["fc_12.bias"]
    probs = torch.softmax(logits, -1)[:, STEPS_TO_DISCARD:]
    if return_intermediates:
        return probs, {"features": feat, "theta": theta, "warped": warped, "fc_9": seq,
                       "l1": l1, "l2": l2, "logits": logits[:, STEPS_TO_DISCARD:]}
    return probs


def ctc_greedy(probs):
    """keras.backend.ctc_decode(greedy=True) + -1 padding (recognition.py:169-184).

    probs: (B,T,K).  Returns int64 (B,T): merged, blank-free labels, padded with -1.
    """
    probs = torch.as_tensor(probs)
    B, T, K = probs.shape
    best = torch.argmax(torch.log(probs + 1e-7), -1).numpy()   # first maximum wins
    out = np.full((B, T), -1, dtype=np.int64)
    for b in range(B):
        n, prev = 0, -1
        for t in range(T):
            c = int(best[b, t])
            if c != K - 1 and c != prev:
                out[b, n] = c
                n += 1
            prev = c
    return out


def labels_to_text(rows, alphabet=ALPHABET):
    """recognition.py:527-534: drop blank and -1, map to characters."""
    blank = len(alphabet)
    return ["".join(alphabet[i] for i in row if i not in (blank, -1)) for row in np.asarray(rows)]


def recognize_crops(weights, crops):
    """crops uint8 (B,31,200) -> list[str]; the /255 scaling follows recognition.py:524."""
    x = np.asarray(crops, dtype="float32") / 255
    return labels_to_text(ctc_greedy(crnn_logits(weights, x)))
