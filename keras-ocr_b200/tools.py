"""Host-side helpers of the hot path (reference keras_ocr/tools.py, hot subset only)."""
import hashlib
import io
import os
import typing

import numpy as np


def read(filepath_or_buffer: typing.Union[str, io.BytesIO, np.ndarray]):
    """tools.read (reference tools.py:19-38): ndarrays pass through; files / buffers are decoded to
    RGB on the host (image decode stays host-side).  URLs need a network and
    are not supported offline."""
    if isinstance(filepath_or_buffer, np.ndarray):
        return filepath_or_buffer
    import cv2

    if hasattr(filepath_or_buffer, "read"):
        data = np.asarray(bytearray(filepath_or_buffer.read()), dtype=np.uint8)
        image = cv2.imdecode(data, cv2.IMREAD_UNCHANGED)
    elif isinstance(filepath_or_buffer, str):
        assert os.path.isfile(filepath_or_buffer), "Could not find image at path: " + filepath_or_buffer
        image = cv2.imread(filepath_or_buffer)
    else:
        raise TypeError(f"cannot read image from {type(filepath_or_buffer)!r}")
    return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)


def resize_plan(shape, max_scale, max_size):
    """The scale and output size tools.resize_image (reference tools.py:378-398) would pick.

    Returns (scale, out_h, out_w).  ``max(shape)`` includes the channel axis, as upstream."""
    longest = max(shape)
    scale = max_size / longest if longest * max_scale > max_size else max_scale
    return scale, int(shape[0] * scale), int(shape[1] * scale)


def adjust_boxes(boxes, scale=1, boxes_format="boxes"):
    """tools.adjust_boxes (reference tools.py:232-260)."""
    if scale == 1:
        return boxes
    if boxes_format == "boxes":
        return np.array(boxes) * scale
    if boxes_format == "lines":
        return [[(np.array(box) * scale, character) for box, character in line] for line in boxes]
    if boxes_format == "predictions":
        return [(word, np.array(box) * scale) for word, box in boxes]
    raise NotImplementedError(f"Unsupported boxes format: {boxes_format}")


def fit(image, width, height, cval=255, mode="letterbox", return_scale=False):
    """tools.fit (reference tools.py:402-452): scale the image to fit ``width`` x ``height`` keeping
    its aspect ratio, then letterbox (pad bottom/right with ``cval``) or crop.  Host-side like the
    reference: it only serves the single-crop ``Recognizer.recognize`` API, which is off the hot path."""
    import cv2

    sx, sy = width / image.shape[1], height / image.shape[0]
    if sx == 1 and sy == 1:
        return (image, 1) if return_scale else image
    if mode not in ("letterbox", "crop"):
        raise NotImplementedError(f"Unsupported mode: {mode}")
    use_width = (sx <= sy) if mode == "letterbox" else (sx >= sy)
    if use_width:
        scale, new_w, new_h = sx, width, sx * image.shape[0]
    else:
        scale, new_h, new_w = sy, height, sy * image.shape[1]
    resized = cv2.resize(image, dsize=(int(new_w), int(new_h)))
    if mode == "letterbox":
        fitted = np.zeros((height, width, 3), dtype="uint8") + cval
        fitted[: resized.shape[0], : resized.shape[1]] = resized[:height, :width]
    else:
        fitted = resized[:height, :width]
    return (fitted, scale) if return_scale else fitted


def read_and_fit(filepath_or_array, width, height, cval=255, mode="letterbox"):
    """tools.read_and_fit (reference tools.py:455-481)."""
    image = read(filepath_or_array) if isinstance(filepath_or_array, str) else filepath_or_array
    return fit(image=image, width=width, height=height, cval=cval, mode=mode)


def sha256sum(filename):
    """tools.sha256sum (reference tools.py:484-492)."""
    h = hashlib.sha256()
    with open(filename, "rb") as f:
        for chunk in iter(lambda: f.read(128 * 1024), b""):
            h.update(chunk)
    return h.hexdigest()


def get_default_cache_dir():
    """tools.get_default_cache_dir (reference tools.py:495-498)."""
    return os.environ.get("KERAS_OCR_CACHE_DIR", os.path.expanduser(os.path.join("~", ".keras-ocr")))


def find_cached(filename, sha256=None, cache_dir=None):
    """Offline half of tools.download_and_verify (reference tools.py:501-530): locate and verify a
    weight file that is already in the cache; there is no network here to download it."""
    path = os.path.join(cache_dir or get_default_cache_dir(), filename)
    assert os.path.isfile(path), (
        f"{path} not found and no network is available to download it; pass weights=<dict|path> instead")
    assert sha256 is None or sha256 == sha256sum(path), "Error occurred verifying sha256."
    return path
