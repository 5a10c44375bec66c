"""Per-step ray batches and checkpoints.

``RayBatchSampler`` replaces the host-side numpy bookkeeping of NeRF/run_nerf.py:304-306,368-407 (shuffled global
ray ids -> pixel coordinates, per-ray camera index, target colours, re-shuffle at the end of an epoch) with one
CUDA kernel over device-resident images.  ``save_checkpoint`` writes the reference's ``.tar`` layout
(run_nerf.py:626-641) so either code base can resume the other's run."""
import os

import torch

from . import _lib


class RayBatchSampler:
    def __init__(self, images, i_train, H, W, N_rand, generator=None):
        """images: [n_images, H, W, 3] float32 CUDA tensor; i_train: indices of the training images."""
        self.images = _lib.f32(images)
        dev = self.images.device
        self.i_train = torch.as_tensor(i_train, dtype=torch.int64, device=dev).contiguous()
        self.H, self.W, self.N_rand = int(H), int(W), int(N_rand)
        self.n_rays = self.i_train.numel() * self.H * self.W
        self.generator = generator
        self.shuffled_ray_idx = None
        self.i_batch = 0
        self.shuffle()

    def shuffle(self, permutation=None):
        """np.random.shuffle(shuffled_ray_idx) (:305, :402), on the device.  ``permutation`` injects one (tests)."""
        dev = self.images.device
        if permutation is None:
            permutation = torch.randperm(self.n_rays, device=dev, generator=self.generator)
        self.shuffled_ray_idx = torch.as_tensor(permutation, dtype=torch.int64, device=dev).contiguous()
        self.i_batch = 0

    def next(self):
        """-> (kps[N,2] int64 (x, y), image_idx[N] int64, target[N,3]) for the next N_rand rays (:368-398)."""
        lib = _lib.load()
        dev = self.images.device
        sl = self.shuffled_ray_idx[self.i_batch:self.i_batch + self.N_rand]
        N = sl.numel()
        kps = torch.empty(N, 2, device=dev, dtype=torch.int64)
        idx = torch.empty(N, device=dev, dtype=torch.int64)
        target = torch.empty(N, 3, device=dev, dtype=torch.float32)
        _lib.check(lib.scnerf_ray_batch(_lib.ptr(sl), N, _lib.ptr(self.images), _lib.ptr(self.i_train),
                                        self.i_train.numel(), self.H, self.W, _lib.ptr(kps), _lib.ptr(idx),
                                        _lib.ptr(target), _lib.stream()), "ray_batch")
        self.i_batch += self.N_rand
        if self.i_batch >= self.n_rays:            # :399-403
            print("Shuffle data after an epoch!")
            self.shuffle()
        return kps, idx, target


def save_checkpoint(path, global_step, render_kwargs_train, optimizer, camera_model=None):
    """The reference's checkpoint dict (run_nerf.py:626-641)."""
    save_dict = {
        "global_step": global_step,
        "network_fn_state_dict": render_kwargs_train["network_fn"].state_dict(),
        "network_fine_state_dict": render_kwargs_train["network_fine"].state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
    }
    if camera_model is not None:
        save_dict["camera_model"] = camera_model.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(save_dict, path)
    return path
