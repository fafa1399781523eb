"""Image / label-map folder dataset in the reference's item format (data/base_dataset.py:44-118:
'resize_and_crop' preprocessing - resize to load_size (labels NEAREST, photos bicubic / bilinear),
random crop_size crop, random horizontal flip while training; labels scaled back to integers with
255 -> label_nc "unknown"; photos normalised to [-1, 1]).  PIL + numpy only."""
import os
import random
import re

import numpy as np
import torch
import torch.utils.data as data
from PIL import Image

IMG_EXTENSIONS = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.tiff', '.webp')


def _natural_key(text):
    return [int(c) if c.isdigit() else c for c in re.split(r'(\d+)', text)]


def _list_images(folder):
    files = [os.path.join(folder, f) for f in os.listdir(folder) if f.lower().endswith(IMG_EXTENSIONS)]
    return sorted(files, key=_natural_key)   # util.natural_sort (base_dataset.py:52-53)


class FolderDataset(data.Dataset):
    def initialize(self, opt):
        self.opt = opt
        self.method = Image.BILINEAR if getattr(opt, "downsampling_method", "bicubic") == "bilinear" else Image.BICUBIC
        self.label_paths = _list_images(opt.label_dir)[:getattr(opt, "max_dataset_size", 1 << 62)]
        self.image_paths = _list_images(opt.image_dir)[:getattr(opt, "max_dataset_size", 1 << 62)]
        if not getattr(opt, "no_pairing_check", False):
            assert len(self.label_paths) == len(self.image_paths), \
                "The #images in %s and %s do not match" % (opt.label_dir, opt.image_dir)
            for a, b in zip(self.label_paths, self.image_paths):
                assert self.paths_match(a, b), "The label-image pair (%s, %s) does not look like a pair" % (a, b)

    @staticmethod
    def paths_match(p1, p2):
        return os.path.splitext(os.path.basename(p1))[0] == os.path.splitext(os.path.basename(p2))[0]

    def __len__(self):
        return len(self.label_paths)

    def _geometry(self):
        opt = self.opt
        x = random.randint(0, max(0, opt.load_size - opt.crop_size))
        y = random.randint(0, max(0, opt.load_size - opt.crop_size))
        flip = random.random() > 0.5
        return x, y, flip and opt.isTrain and not getattr(opt, "no_flip", False)

    def _apply(self, img, method, geom):
        opt = self.opt
        x, y, flip = geom
        img = img.resize((opt.load_size, opt.load_size), method)
        img = img.crop((x, y, x + opt.crop_size, y + opt.crop_size))
        return img.transpose(Image.FLIP_LEFT_RIGHT) if flip else img

    def __getitem__(self, index):
        geom = self._geometry()
        label = self._apply(Image.open(self.label_paths[index]), Image.NEAREST, geom)
        lab = torch.from_numpy(np.asarray(label, dtype=np.uint8).copy()).float()
        if lab.dim() == 3:
            lab = lab[..., 0]
        lab = lab.unsqueeze(0)                       # ToTensor() * 255 of a mode-L / P image
        lab[lab == 255] = self.opt.label_nc          # 'unknown' (base_dataset.py:99)
        image = self._apply(Image.open(self.image_paths[index]).convert('RGB'), self.method, geom)
        img = torch.from_numpy(np.asarray(image, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        img = (img - 0.5) / 0.5
        return {'label': lab, 'image': img, 'path': self.image_paths[index]}


class CelebAMaskHQDataset(FolderDataset):
    """data/celebamaskhq_dataset.py without the identity-file guided sampling (that needs the
    dataset's identity CSV; guided models take `guiding_image` / `guiding_label` from the caller)."""


def find_dataset_using_name(name):
    table = {"celebamaskhq": CelebAMaskHQDataset, "celeba": FolderDataset, "folder": FolderDataset}
    key = name.replace('_', '').lower()
    if key not in table:
        raise ValueError("dataset_mode %r is not available on the B200 path (have: %s)" % (name, sorted(table)))
    return table[key]
