"""Device time of the image-module kernels at the Weizmann shapes (C4: 25 x 25 = 625 frames of 3 x 64 x 64, n_kernels 64,
z_dim 256) next to the torch / cuDNN modules they replace, forward + backward of ImageEncoder -> ImageDecoder and
layer by layer.  CUDA events on the current stream, 3 warm-up + 10 timed passes.  `python tools/time_conv.py [frames]`"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_dmm_b200.models.common as common  # noqa: E402


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 625
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    enc = common.ImageEncoder(256).to(dev).train()
    dec = common.ImageDecoder(256).to(dev).train()
    x = torch.rand(frames, 3, 64, 64, device=dev)
    tgt = (torch.rand(frames, 3, 64, 64, device=dev) > 0.5).float()

    def step():
        for m in (enc, dec):
            m.zero_grad(set_to_none=True)
        mean, std = enc(x)
        (probs,) = dec(mean + 0.1 * std)
        F.binary_cross_entropy(probs, tgt, reduction='sum').backward()

    out = {'frames': frames}
    for label, flags in (('kernels', {'conv': True, 'dense': True}), ('torch_cudnn_fp32', {'conv': False, 'dense': False})):
        common.IMAGE_KERNELS.update(flags)
        out[label + '_enc_dec_fwd_bwd_ms'] = timed(step)
    common.IMAGE_KERNELS.update({'conv': True, 'dense': True})
    # layer by layer: forward, backward (input gradient + weight gradient + bias gradient)
    layers = [('conv_stack.0', enc.conv_stack[0], (frames, 3, 64, 64)), ('conv_stack.1', enc.conv_stack[1], (frames, 16, 32, 32)),
              ('conv_stack.2', enc.conv_stack[2], (frames, 32, 16, 16)), ('deconv_stack.0', dec.deconv_stack[0], (frames, 64, 8, 8)),
              ('deconv_stack.1', dec.deconv_stack[1], (frames, 32, 16, 16)), ('deconv_stack.2', dec.deconv_stack[2], (frames, 16, 32, 32))]
    per = {}
    for name, blk, shape in layers:
        xi = torch.randn(*shape, device=dev, requires_grad=True)
        row = {}
        for label, on in (('kernels', True), ('torch', False)):
            common.IMAGE_KERNELS['conv'] = on
            y = blk(xi)
            dy = torch.randn_like(y)
            row[label + '_fwd_ms'] = timed(lambda: blk(xi))

            def fb():
                blk.zero_grad(set_to_none=True)
                xi.grad = None
                blk(xi).backward(dy)
            row[label + '_fwd_bwd_ms'] = timed(fb)
        layer = blk.conv if hasattr(blk, 'conv') else blk.deconv
        k = layer.kernel_size[0]
        if hasattr(blk, 'conv'):
            macs = frames * layer.out_channels * (shape[2] // 2) * (shape[3] // 2) * layer.in_channels * k * k
        else:
            macs = frames * layer.in_channels * shape[2] * shape[3] * layer.out_channels * k * k
        row['fwd_gflop'] = 2 * macs / 1e9
        row['kernels_fwd_bwd_tflops'] = 3 * 2 * macs / (row['kernels_fwd_bwd_ms'] * 1e-3) / 1e12
        per[name] = row
    common.IMAGE_KERNELS['conv'] = True
    out['layers'] = per
    print(json.dumps(out))


if __name__ == '__main__':
    main()
