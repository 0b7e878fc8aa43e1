"""The pseudo-ground-truth records that connect the two reference drivers: run_reconstruction.py --export-pseudogt writes
one `<idx>.npz` per training image (:506-610) and main.py's datasets read them back (data/abstract_dataset.py:68-107).
Host-side only (numpy / torch on the CPU); the tensors come from the CUDA renderer.

Record format (kept byte-compatible so that either side can be swapped for the reference's):
    np.savez_compressed(<cache>/<dataset>/pseudogt_<R>x<R>/<idx>, data={
        'mesh':          float32 [3, 32, 32]   predicted displacement map,
        'texture':       float16 [3, R, R]     image projected onto the UV map, masked by texel visibility,
        'texture_alpha': float16 [1, R, R]     projected alpha, same mask,
        'image':         float16 [3|4, h, w]   the Inception-resolution image in [-1, 1]})
i.e. ONE pickled dict of torch tensors under the key 'data' (np.load(..., allow_pickle=True)['data'].item()).
`poses_metadata.npz` uses the same convention with keys scale / translation / rotation / path (:611-620).
"""
import os

import numpy as np
import torch
import torch.nn.functional as F


def pseudo_gt_dir(cache_dir, resolution):
    return os.path.join(cache_dir, f'pseudogt_{resolution}x{resolution}')


def visibility_to_mask(visibility, resolution):
    """Texel-visibility mask at the pseudo-GT resolution from d(render)/d(texture) (run_reconstruction.py:571-583):
    bilinear resize, then "any channel received gradient".  visibility [B,3,Th,Tw] -> [B,R,R,1] float {0,1}."""
    m = F.interpolate(visibility, resolution, mode='bilinear', align_corners=False).permute(0, 2, 3, 1)
    return (m > 0).any(dim=3, keepdim=True).float()


def make_record(mesh_map, inverse_tex, inverse_alpha, inception_image):
    """One sample's record from NCHW tensors (:586-603): textures and image stored as fp16, the displacement map as is."""
    return {
        'mesh': mesh_map.detach().cpu().clone(),
        'texture': inverse_tex.detach().half().cpu().clone(),
        'texture_alpha': inverse_alpha.detach().half().cpu().clone(),
        'image': inception_image.detach().half().cpu().clone(),
    }


def save_pseudo_gt(directory, idx, record):
    os.makedirs(directory, exist_ok=True)
    np.savez_compressed(os.path.join(directory, f'{idx}'), data=record)


def load_pseudo_ground_truth(cache_dir, texture_resolution, idx):
    """What AbstractDataset.load_pseudo_ground_truth returns (abstract_dataset.py:68-81): image rescaled to [0, 1] and cut
    to RGB, textures widened back to fp32, the displacement map untouched."""
    data = np.load(os.path.join(pseudo_gt_dir(cache_dir, texture_resolution), f'{idx}.npz'), allow_pickle=True)['data'].item()
    return {
        'image': data['image'][:3].float() / 2 + 0.5,
        'texture': data['texture'].float(),
        'texture_alpha': data['texture_alpha'].float(),
        'mesh': data['mesh'],
    }


def mirror_tex(tr):
    """"Virtually" mirror a texture / displacement map [C,H,W] (abstract_dataset.py:99-107): flip along u and shift u by
    half a turn, which is what re-projecting the mirrored photograph would give."""
    tr = torch.flip(tr, dims=(2,))
    tr = torch.cat((tr, tr), dim=2)
    q = tr.shape[2] // 4
    return tr[:, :, q:-q]


def save_poses_metadata(cache_dir, scale, translation, rotation, paths):
    os.makedirs(cache_dir, exist_ok=True)
    np.savez_compressed(os.path.join(cache_dir, 'poses_metadata'),
                        data={'scale': scale, 'translation': translation, 'rotation': rotation, 'path': list(paths)})


def load_poses_metadata(cache_dir):
    return np.load(os.path.join(cache_dir, 'poses_metadata.npz'), allow_pickle=True)['data'].item()
