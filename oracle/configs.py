"""Named test/bench configurations shared by make_golden.py, tests/ and bench.py
(TEST INFRASTRUCTURE).  Hyper-parameters come from the reference's shipped configs:
  ccdm/params.yml:69-75 + builder.py:26-28         (CCDM 3-D UNet)
  latentdiffusion/configs/latent-diffusion/ruijin-ldm_from_controlnet_ae.yaml:17-40  (LDM _ae)
  latentdiffusion/configs/latent-diffusion/ruijin-ldm_from_controlnet.yaml:17-40     (LDM pixel)
"""

CCDM_PARAMS_YML = dict(base_channels=64, channel_mult=[1, 2, 2, 4, 5], attention_resolutions=[32, 16, 8],
                       num_heads=1, num_head_channels=32, softmax_output=True)

CCDM_TINY = dict(base_channels=32, channel_mult=[1, 2], attention_resolutions=[2],
                 num_heads=1, num_head_channels=32, softmax_output=True)

LDM_AE = dict(dims=2, image_size=512, in_channels=8, out_channels=4, model_channels=160,
              attention_resolutions=[8, 4, 2], num_res_blocks=2, channel_mult=[1, 2, 4, 4, 5], num_head_channels=32)

LDM_PIXEL = dict(dims=2, image_size=512, in_channels=3, out_channels=1, model_channels=128,
                 attention_resolutions=[32, 16, 8], num_res_blocks=2, channel_mult=[1, 2, 4, 4, 5],
                 num_head_channels=32)

LDM_TINY = dict(dims=2, image_size=16, in_channels=8, out_channels=4, model_channels=32,
                attention_resolutions=[2], num_res_blocks=1, channel_mult=[1, 2], num_head_channels=32)

LDM_TINY_XATTN = dict(LDM_TINY, use_spatial_transformer=True, transformer_depth=1, context_dim=64)

LDM_SCHEDULE = dict(timesteps=1000, linear_start=0.0015, linear_end=0.0195)
