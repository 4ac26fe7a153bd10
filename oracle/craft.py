"""Oracle: CRAFT detector forward pass, fp32 on the CPU (torch functional ops).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the graph assembled by
``build_keras_model`` (reference keras_ocr/detection.py:353-424) with the layer
recipes of ``make_vgg_block`` (87-103) and ``upconv`` (65-84).  The weights are a
flat dict using the reference's own PyTorch key names (the names
``load_torch_weights`` maps, detection.py:428-468): conv ``<name>.weight`` OIHW +
``<name>.bias``; batch-norm ``<name>.{weight,bias,running_mean,running_var}``.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # detection.py:69-71, 80-82, 95-97 (explicit epsilon=1e-5)

# (conv name, bn name, maxpool-after) for the 12 encoder convs, detection.py:312-324.
ENCODER = [
    ("basenet.slice1.0", "basenet.slice1.1", False),
    ("basenet.slice1.3", "basenet.slice1.4", True),
    ("basenet.slice1.7", "basenet.slice1.8", False),
    ("basenet.slice1.10", "basenet.slice1.11", True),   # relu output = tap s1 (slice1.12)
    ("basenet.slice2.14", "basenet.slice2.15", False),
    ("basenet.slice2.17", "basenet.slice2.18", False),  # relu output = tap s2 (slice2.19)
    ("basenet.slice3.20", "basenet.slice3.21", True),
    ("basenet.slice3.24", "basenet.slice3.25", False),
    ("basenet.slice3.27", "basenet.slice3.28", False),  # relu output = tap s3 (slice3.29)
    ("basenet.slice4.30", "basenet.slice4.31", True),
    ("basenet.slice4.34", "basenet.slice4.35", False),
    ("basenet.slice4.37", "basenet.slice4.38", False),  # BN output (NO relu) = tap s4
]


def _conv(w, x, name, padding=0, dilation=1):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], padding=padding, dilation=dilation)


def _bn(w, x, name):
    return F.batch_norm(
        x,
        w[name + ".running_mean"],
        w[name + ".running_var"],
        w[name + ".weight"],
        w[name + ".bias"],
        training=False,
        eps=BN_EPS,
    )


def _upconv(w, x, n):
    """detection.py:65-84: 1x1 conv + BN + ReLU, 3x3 conv + BN + ReLU."""
    x = F.relu(_bn(w, _conv(w, x, f"upconv{n}.conv.0"), f"upconv{n}.conv.1"))
    x = F.relu(_bn(w, _conv(w, x, f"upconv{n}.conv.3", padding=1), f"upconv{n}.conv.4"))
    return x


def _upsample_like(src, target):
    """UpsampleLike, detection.py:290-309: resize_bilinear(half_pixel_centers=True)."""
    return F.interpolate(src, size=target.shape[2:], mode="bilinear", align_corners=False)


def craft_forward(weights, x, return_taps=False):
    """x: (N,3,H,W) float32, already normalised by compute_input.

    Returns the score maps as (N, H/2, W/2, 2) float32 (channel 0 = text/region,
    channel 1 = link/affinity), like the Keras model output (detection.py:408-413).
    """
    w = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in weights.items()}
    taps = {}
    h = x
    for conv_name, bn_name, pool in ENCODER:
        h = _bn(w, _conv(w, h, conv_name, padding=1), bn_name)
        if conv_name == "basenet.slice4.37":
            taps["s4"] = h            # detection.py:333 taps the BN layer, not the ReLU
            break
        h = F.relu(h)
        if conv_name == "basenet.slice1.10":
            taps["s1"] = h
        elif conv_name == "basenet.slice2.17":
            taps["s2"] = h
        elif conv_name == "basenet.slice3.27":
            taps["s3"] = h
        if pool:
            h = F.max_pool2d(h, 2, 2)
    s1, s2, s3, s4 = taps["s1"], taps["s2"], taps["s3"], taps["s4"]

    # slice5, detection.py:365-378: maxpool 3x3/1 "same", dilated conv, 1x1 conv (no BN, no act).
    s5 = F.max_pool2d(s4, 3, 1, 1)
    s5 = _conv(w, s5, "basenet.slice5.1", padding=6, dilation=6)
    s5 = _conv(w, s5, "basenet.slice5.2")

    y = torch.cat([s5, s4], 1)                       # detection.py:380
    y = _upconv(w, y, 1)
    y = torch.cat([_upsample_like(y, s3), s3], 1)    # 382-383
    y = _upconv(w, y, 2)
    y = torch.cat([_upsample_like(y, s2), s2], 1)    # 385-386
    y = _upconv(w, y, 3)
    y = torch.cat([_upsample_like(y, s1), s1], 1)    # 388-389
    feat = _upconv(w, y, 4)

    # conv_cls head, detection.py:392-410; last layer linear for the vgg backbone (411-412).
    y = F.relu(_conv(w, feat, "conv_cls.0", padding=1))
    y = F.relu(_conv(w, y, "conv_cls.2", padding=1))
    y = F.relu(_conv(w, y, "conv_cls.4", padding=1))
    y = F.relu(_conv(w, y, "conv_cls.6"))
    y = _conv(w, y, "conv_cls.8")
    out = y.permute(0, 2, 3, 1).contiguous()
    if return_taps:
        taps["feature"] = feat
        return out, taps
    return out
