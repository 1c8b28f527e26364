"""guidegen-b200: B200-native (sm_100a) implementation of GuideGen's reverse-diffusion denoising
step -- the CCDM categorical mask sampler and the LDM DDIM CT sampler -- behind the reference's
Python class surface.  All arithmetic runs in libguidegen_sm100.so (include/guidegen_sm100.h);
there is no PyTorch / CPU fallback.  See DESIGN.md."""
__version__ = "0.1.0"
