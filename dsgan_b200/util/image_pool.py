"""History buffer of generated images for the discriminator (reference: DSGAN/util/image_pool.py:4-32).
Host-side logic on device tensors: the first `pool_size` images pass through; afterwards each image is swapped
with a random stored one with probability 0.5 (python `random`, like the reference)."""
import random

import torch


class ImagePool:
    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        if self.num_imgs + images.shape[0] <= self.pool_size:
            # fast path (identical result): every image is stored and returned unchanged
            for img in images:
                self.images.append(img.detach().unsqueeze(0).clone())
            self.num_imgs += images.shape[0]
            return images.detach()
        out = []
        for img in images:
            img = img.detach().unsqueeze(0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(img.clone())
                out.append(img)
            elif random.uniform(0, 1) > 0.5:
                idx = random.randint(0, self.pool_size - 1)
                out.append(self.images[idx].clone())
                self.images[idx] = img.clone()
            else:
                out.append(img)
        return torch.cat(out, 0)
