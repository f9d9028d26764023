"""blurred_gan_b200 -- B200-native (sm_100a) hot path of lebrice/blurred-GAN.

Drop-in modules, same names and call signatures as the reference files they replace:

    blurred_gan_b200.pytorch.gaussianblur   <- pytorch/gaussianblur.py   (GaussianBlur(std))
    blurred_gan_b200.gaussian_blur          <- gaussian_blur.py          (blur_images, gaussian_blur, GaussianBlur2D)
    blurred_gan_b200.sliced_wasserstein     <- sliced_wasserstein.py     (API, free functions)

All arithmetic runs in hand-written CUDA kernels in libbgan_b200.so (C ABI: include/bgan.h),
reached through the PyTorch custom ops in blurred_gan_b200.ops.  There is no CPU fallback.
"""
__version__ = "0.1.0"
