"""Image loading for the extraction loop: PIL decode -> transform chain -> normalised NCHW fp32 tensors.

Mirrors ``dirtorch/utils/pytorch_loader.get_loader`` (pytorch_loader.py:11-73) and the part of
``dirtorch/utils/transforms.create`` the evaluation CLIs use (transforms.py:11-37: the chain always ends in
ToTensor + Normalize(mean, std); the deterministic test-time transforms ``Scale`` :133-185, ``Pad`` :47-77,
``PadSquare`` :79-104, ``CenterCrop`` :309-322, ``Identity`` :40-44).  CPU-side I/O, not part of the GPU hot path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.utils.data as data
from PIL import Image


class Scale:
    """Rescale a PIL image: int = smallest side, float = factor (rounded as int(0.5 + s*w)), tuple = (w, h)."""

    def __init__(self, size, interpolation=Image.BILINEAR, largest=False, can_upscale=True, can_downscale=True):
        assert isinstance(size, (float, int)) or len(size) == 2
        if isinstance(size, float):
            assert 0 < size <= 4, "bad float size, cannot be outside of range ]0,4]"
        self.size, self.interpolation, self.largest = size, interpolation, largest
        self.can_upscale, self.can_downscale = can_upscale, can_downscale

    def get_params(self, imsize):
        w, h = imsize
        if isinstance(self.size, int):
            smaller = (lambda a, b: a >= b) if self.largest else (lambda a, b: a <= b)
            if (smaller(w, h) and w == self.size) or (smaller(h, w) and h == self.size):
                return w, h
            if smaller(w, h):
                return self.size, int(0.5 + self.size * h / w)
            return int(0.5 + self.size * w / h), self.size
        if isinstance(self.size, float):
            return int(0.5 + self.size * w), int(0.5 + self.size * h)
        return tuple(self.size)

    def __call__(self, img):
        size2 = self.get_params(img.size)
        if size2 != img.size:
            a1, a2 = img.size, size2
            if (self.can_upscale and min(a1) < min(a2)) or (self.can_downscale and min(a1) > min(a2)):
                img = img.resize(size2, self.interpolation)
        return img


def _rgb_fill(color):
    assert len(color) == 3
    return tuple(color) if all(isinstance(c, int) for c in color) else tuple(int(255 * c) for c in color)


class Identity:
    """transforms.py:40-44."""

    def __call__(self, img):
        return img


class Pad:
    """Pad the SHORTEST side up to `size` with a constant colour, image centred (transforms.py:47-77)."""

    def __init__(self, size, color=(127, 127, 127)):
        self.size, self.color = size, _rgb_fill(color)

    def __call__(self, img):
        w, h = img.size
        target = (w, max(h, self.size)) if w >= h else (max(w, self.size), h)
        if target == img.size:
            return img
        canvas = Image.new("RGB", target, self.color)
        canvas.paste(img, ((target[0] - w) // 2, (target[1] - h) // 2))
        return canvas


class PadSquare:
    """Pad (or centre-crop, when `size` is smaller) to size x size; size=None -> the larger side (transforms.py:79-104)."""

    def __init__(self, size=None, color=(127, 127, 127)):
        self.size, self.color = size, _rgb_fill(color)

    def __call__(self, img):
        w, h = img.size
        s = self.size or max(w, h)
        if (s, s) == img.size:
            return img
        canvas = Image.new("RGB", (s, s), self.color)
        canvas.paste(img, ((s - w) // 2, (s - h) // 2))
        return canvas


class CenterCrop:
    """Crop `size` = int or (h, w) around the centre, offsets rounded as int(0.5 + d/2) (transforms.py:309-322);
    `padding` > 0 first adds a black border (RandomCrop.__call__, transforms.py:286-305)."""

    def __init__(self, size, padding=0):
        self.size = (int(size), int(size)) if isinstance(size, int) else tuple(size)
        self.padding = padding

    def __call__(self, img):
        if self.padding > 0:
            from PIL import ImageOps
            img = ImageOps.expand(img, border=self.padding, fill=0)
        w, h = img.size
        th, tw = self.size
        x, y = int(0.5 + (w - tw) / 2.0), int(0.5 + (h - th) / 2.0)
        return img.crop((x, y, x + tw, y + th))


class ToTensor:
    def __call__(self, img):
        a = np.array(img, dtype=np.uint8)
        return torch.from_numpy(a).permute(2, 0, 1).to(torch.float32).div_(255.0)


class Normalize:
    def __init__(self, mean, std):
        self.mean = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
        self.std = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)

    def __call__(self, t):
        return (t - self.mean) / self.std


class Compose:
    def __init__(self, trfs):
        self.trfs = list(trfs)

    def __call__(self, x):
        for t in self.trfs:
            x = t(x)
        return x


def create_transforms(cmd_line, to_tensor=False, **vars):
    """Comma-separated transform list -> callable; ToTensor + Normalize(mean, std) appended like the reference."""
    if to_tensor:
        if not cmd_line:
            cmd_line = "ToTensor(), Normalize(mean=mean, std=std)"
        elif "ToTensor" not in cmd_line:
            cmd_line += ", ToTensor(), Normalize(mean=mean, std=std)"
    assert isinstance(cmd_line, str)
    env = {"Scale": Scale, "Identity": Identity, "Pad": Pad, "PadSquare": PadSquare, "CenterCrop": CenterCrop,
           "ToTensor": ToTensor, "Normalize": Normalize, "Image": Image}
    env.update(vars)
    import re
    unknown = [n for n in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", cmd_line) if n not in env]
    if unknown:
        raise SyntaxError("Cannot interpret this transform list: %s\nReason: unsupported transform(s) %s - the evaluation "
                          "path implements the deterministic test-time transforms %s (the reference's random training "
                          "augmentations are out of scope)" % (cmd_line, sorted(set(unknown)), sorted(k for k in env if k[0].isupper() and k != "Image")))
    try:
        return Compose(eval("[%s]" % cmd_line, {"__builtins__": {}}, env))
    except Exception as e:
        raise SyntaxError("Cannot interpret this transform list: %s\nReason: %s" % (cmd_line, e))


class _Items(data.Dataset):
    def __init__(self, dataset, transform, output):
        self.dataset, self.transform, self.output = dataset, transform, output

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, i):
        item = {"img_key": self.dataset.get_key(i), "img_filename": self.dataset.get_filename(i)}
        item["img"] = self.transform(self.dataset.get_image(i))
        return [item[o] for o in self.output]


def get_loader(dataset, trf_chain, iscuda, preprocess={}, output=("img", "label"), batch_size=None, threads=1,
               shuffle=False, **_useless_kw):
    """Iterable of ``[img_batch]`` (pytorch_loader.py:11-73); always batched (the reference returns an un-batched
    dataset for threads == 1, which its own extraction loop cannot consume)."""
    trf = create_transforms(trf_chain, to_tensor=True, **preprocess)
    items = _Items(dataset, trf, list(output))
    return data.DataLoader(items, batch_size=batch_size or 1, shuffle=shuffle, num_workers=max(0, threads if threads > 1 else 0),
                           pin_memory=bool(iscuda))
