"""CPU simulation of the NUMERICS of the GPU extraction path (TEST INFRASTRUCTURE): the same network as
oracle/dir_oracle.py, with every rounding the kernels perform - fp16 input, fp16 weights, fp32 accumulation, BatchNorm as
an fp32 (scale, shift) epilogue, fp16 activation stores, BN-scaled fp16 weights for the fused projection shortcut, fp32
head.  It answers "does this precision design meet the 1e-3 descriptor bar, and with what margin" without a GPU
(DESIGN.md section 2); accumulation ORDER inside a dot product is the only thing it does not reproduce."""
import torch
import torch.nn.functional as F

from oracle import dir_oracle as O


def h(x):
    """Round to fp16 and come back (what a store to an fp16 tensor + reload does)."""
    return x.to(torch.float16).to(torch.float32)


def _fold(sd, name):
    s = sd[name + ".weight"] / torch.sqrt(sd[name + ".running_var"] + O.BN_EPS)       # net.cu: pack_conv
    return s, sd[name + ".bias"] - sd[name + ".running_mean"] * s


def _conv(x16, w, scale, shift, stride=1, padding=0, res16=None, relu=True):
    y = F.conv2d(x16, h(w), None, stride=stride, padding=padding)                     # fp16 operands, fp32 accumulate
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)                        # epilogue in fp32
    if res16 is not None:
        y = y + res16
    return h(F.relu(y) if relu else y)


@torch.no_grad()
def extract(x, sd, arch="resnet50_rmac", fuse_shortcut=True, **head_kw):
    blocks = O.BLOCKS[arch.split("_")[0]]
    s, b = _fold(sd, "bn1")
    t = _conv(h(x), sd["conv1.weight"], s, b, stride=2, padding=3)
    t = F.max_pool2d(t, kernel_size=3, stride=2, padding=1)
    for li, nblk in enumerate(blocks, start=1):
        for bi in range(nblk):
            p = "layer%d.%d." % (li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            s1, b1 = _fold(sd, p + "bn1")
            s2, b2 = _fold(sd, p + "bn2")
            s3, b3 = _fold(sd, p + "bn3")
            t1 = _conv(t, sd[p + "conv1.weight"], s1, b1)
            t2 = _conv(t1, sd[p + "conv2.weight"], s2, b2, stride=stride, padding=1)
            if bi == 0:
                sdn, bdn = _fold(sd, p + "downsample.1")
                if fuse_shortcut:      # one GEMM over K = [t2 | x] with BN-scaled fp16 weights (net.cu: wcat)
                    y = F.conv2d(t2, h(sd[p + "conv3.weight"] * s3.view(-1, 1, 1, 1))) + \
                        F.conv2d(t, h(sd[p + "downsample.0.weight"] * sdn.view(-1, 1, 1, 1)), stride=stride)
                    t = h(F.relu(y + (b3 + bdn).view(1, -1, 1, 1)))
                else:
                    r = _conv(t, sd[p + "downsample.0.weight"], sdn, bdn, stride=stride, relu=False)
                    t = _conv(t2, sd[p + "conv3.weight"], s3, b3, res16=r)
            else:
                t = _conv(t2, sd[p + "conv3.weight"], s3, b3, res16=t)
    return O.head(t, sd, **head_kw)                                                   # fp32 head on the fp16 map
