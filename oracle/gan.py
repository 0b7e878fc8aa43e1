"""CPU oracle for the conv-GAN path: functional restatement of the reference's generator / discriminators on a
state dict with the reference's key names, in plain torch (F.conv2d, F.batch_norm ...).

TEST INFRASTRUCTURE ONLY.  Parity status: PINNED — tests/test_gan_oracle.py reproduces, from seeded weights,
the golden vectors that tests/golden/make_golden_gan.py produced by running the reference's own modules
(models/gan.py, utils/losses.py imported unmodified from /root/reference/code), forward, losses and gradients.

Restates (under /root/reference/code): models/gan.py:9-20 (positional_encoding), :23-121 (MeshDiscriminator),
:123-233 (TextureDiscriminator), :235-260 (MultiScaleDiscriminator), :264-286 (ConditionalBatchNorm2d), :288-312
(ResBlockUp), :314-426 (Generator); torch.nn.utils.spectral_norm's forward (one power iteration in training mode);
utils/losses.py:21-120 (GANLoss, hinge); main.py:476-521 (ModelWrapper.forward 'g' / 'd').
Buffers (spectral-norm u/v, batch-norm running statistics) in `sd` are updated in place in training mode, like
the modules do."""
import math

import torch
import torch.nn.functional as F


def positional_encoding(Ny, Nx):
    sym = (Nx == Ny // 2)
    ty = torch.arange(Ny, dtype=torch.float64) * (math.pi / Ny)
    tx = -math.pi + torch.arange(Ny, dtype=torch.float64) * (2 * math.pi / Ny)
    X, Y = ty.view(Ny, 1).expand(Ny, Ny), tx.view(1, Ny).expand(Ny, Ny)
    r = torch.stack((X.cos(), X.sin(), Y.cos(), Y.sin()))
    return (r[:, :, Ny // 4: Ny - Ny // 4] if sym else r).float()


def circpad(x, a):
    return torch.cat((x[..., -a:], x, x[..., :a]), dim=3)


def sn_weight(sd, name, training, eps=1e-12):
    """weight of a spectral-normalised layer `name` (keys name.weight_orig / _u / _v)."""
    w, u, v = sd[name + ".weight_orig"], sd[name + ".weight_u"], sd[name + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
    sigma = torch.dot(u.clone(), torch.mv(wm, v.clone()))
    return w / sigma


def conv(x, w, b=None, stride=1, pad_y=0):
    return F.conv2d(x, w, b, stride=stride, padding=(pad_y, 0))


def cbn(sd, name, x, z, training, kind="syncbatch"):
    if kind in ("syncbatch", "batch"):
        x = F.batch_norm(x, sd[name + ".norm.running_mean"], sd[name + ".norm.running_var"], None, None, training, 0.1, 1e-5)
        if training:
            sd[name + ".norm.num_batches_tracked"] += 1
    elif kind == "instance":
        x = F.instance_norm(x)
    g = F.linear(z, sd[name + ".fc_gamma.weight"], sd[name + ".fc_gamma.bias"])[:, :, None, None]
    b = F.linear(z, sd[name + ".fc_beta.weight"], sd[name + ".fc_beta.bias"])[:, :, None, None]
    return x * (1 + g) + b


def resblock(sd, name, x, z, pad, training, kind):
    skip = conv(x, sn_weight(sd, name + ".shortcut", training)) if (name + ".shortcut.weight_orig") in sd else x
    h = F.leaky_relu(cbn(sd, name + ".norm1", conv(pad(x, 1), sn_weight(sd, name + ".conv1", training), pad_y=1), z, training, kind), 0.2)
    h = F.leaky_relu(cbn(sd, name + ".norm2", conv(pad(h, 1), sn_weight(sd, name + ".conv2", training), pad_y=1), z, training, kind), 0.2)
    return h + skip


def symmetrize(x):
    xf = torch.flip(x, (3,))
    w = xf.shape[3]
    return torch.cat((xf[..., w // 2:], x, xf[..., :w // 2]), dim=-1)


def adjust_poles(t):
    top = t[:, :, :1].mean(dim=3, keepdim=True).expand(-1, -1, -1, t.shape[3])
    bot = t[:, :, -1:].mean(dim=3, keepdim=True).expand(-1, -1, -1, t.shape[3])
    return torch.cat((top, t[:, :, 1:-1], bot), dim=2)


def generator(sd, args, z, c, training=True):
    """Generator.forward (symmetric, mesh head, class-conditional)  -> (tex [B,3,R,R], mesh [B,3,32,32])."""
    pad = lambda x, a: F.pad(x, (a, a, 0, 0), mode="replicate")
    up = lambda x: F.interpolate(x, scale_factor=2, mode="nearest")
    kind = args.norm_g
    z = torch.cat((z, F.embedding(c[:, 0], sd["emb_class.weight"])), dim=1)
    x = F.linear(z, sd["fc.weight"], sd["fc.bias"]).view(z.shape[0], -1, 8, 4)
    x = up(resblock(sd, "blk1", x, z, pad, training, kind))
    x = up(resblock(sd, "blk2", x, z, pad, training, kind))
    t = x
    for name in ("blk3a", "blk3b", "blk3c"):
        if (name + ".conv1.weight_orig") in sd:
            t = up(resblock(sd, name, t, z, pad, training, kind))
    t = up(resblock(sd, "blk4", t, z, pad, training, kind))
    t = up(resblock(sd, "blk5", t, z, pad, training, kind))
    t = F.leaky_relu(resblock(sd, "blk6", t, z, pad, training, kind), 0.2)
    tex = torch.tanh(conv(pad(t, 2), sd["conv_final.weight"], sd["conv_final.bias"], pad_y=2))
    m = F.leaky_relu(resblock(sd, "blk3_mesh", x, z, pad, training, kind), 0.2)
    mesh = adjust_poles(conv(pad(m, 2), sd["conv_mesh.weight"], sd["conv_mesh.bias"], pad_y=2))
    return symmetrize(tex), symmetrize(mesh)


def _project(sd, pre, y, feat, c):
    emb = F.embedding(c[:, 0], sd[pre + "projector.weight"])
    return y + (feat * emb[:, :, None, None]).sum(dim=1, keepdim=True)


def texture_discriminator(sd, pre, args, x, c, downsample, training):
    stride_first = (downsample == 1 and args.texture_resolution >= 512) or args.texture_resolution >= 1024
    if downsample > 1:
        x = F.avg_pool2d(x, downsample)
    mask = F.avg_pool2d(x[:, 3:4], 16 if stride_first else 8).detach()
    x = torch.cat((x, positional_encoding(x.shape[2], x.shape[3]).to(x).unsqueeze(0).expand(x.shape[0], -1, -1, -1)), dim=1)
    w = lambda n: sn_weight(sd, pre + n, training)
    b = lambda n: sd.get(pre + n + ".bias")
    if stride_first:
        x = F.leaky_relu(conv(circpad(x, 1), w("conv1"), b("conv1"), 2, 1), 0.2)
    else:
        x = F.leaky_relu(conv(circpad(x, 2), w("conv1"), b("conv1"), 1, 2), 0.2)
    for n in ("conv2", "conv3", "conv4"):
        x = F.leaky_relu(conv(circpad(x, 1), w(n), b(n), 2, 1), 0.2)
    y = conv(circpad(x, 2), w("conv5"), b("conv5"), 1, 2)
    return _project(sd, pre, y, x, c), mask


def mesh_discriminator(sd, pre, args, tex, mesh, c, training):
    x = F.avg_pool2d(tex, tex.shape[2] // mesh.shape[2])
    x = torch.cat((x, mesh, positional_encoding(x.shape[2], x.shape[3]).to(x).unsqueeze(0).expand(x.shape[0], -1, -1, -1)), dim=1)
    mask = F.avg_pool2d(x[:, 3:4], 4).detach()
    w = lambda n: sn_weight(sd, pre + n, training)
    b = lambda n: sd.get(pre + n + ".bias")
    x = F.leaky_relu(conv(circpad(x, 2), w("conv1"), b("conv1"), 1, 2), 0.2)
    x = F.leaky_relu(conv(circpad(x, 1), w("conv2"), b("conv2"), 2, 1), 0.2)
    x = F.leaky_relu(conv(circpad(x, 1), w("conv3"), b("conv3"), 2, 1), 0.2)
    y = conv(circpad(x, 2), w("conv4"), b("conv4"), 1, 2)
    return _project(sd, pre, y, x, c), mask


def discriminator(sd, args, x, mesh, c, training=True):
    d1, m1 = texture_discriminator(sd, "d1.", args, x, c, 1, training)
    d2, m2 = mesh_discriminator(sd, "d2.", args, x, mesh, c, training)
    outs, masks = [d1, d2], [m1, m2]
    if args.num_discriminators == 3:
        d3, m3 = texture_discriminator(sd, "d3.", args, x, c, 4, training)
        outs.append(d3)
        masks.append(m3)
    return outs, masks


def hinge(preds, target_is_real, for_discriminator, masks, weight=None):
    """GANLoss('hinge').__call__ on lists (utils/losses.py:81-120)."""
    total = 0
    for i, (p, m) in enumerate(zip(preds, masks)):
        if for_discriminator:
            v = torch.clamp((p if target_is_real else -p) - 1, max=0)
        else:
            v = p
        per = (v * m).flatten(1).sum(1) / m.flatten(1).sum(1)
        total = total - per.mean() * (1 if weight is None else weight[i])
    return total / (len(preds) if weight is None else sum(weight))


def g_loss(sg, sd_, args, z, c, alpha, training=True):
    tex, mesh = generator(sg, args, z, c, training)
    out, mask = discriminator(sd_, args, torch.cat((tex * alpha, alpha), 1), mesh, c, training)
    return hinge(out, True, False, mask), tex, mesh, out, mask


def d_loss(sg, sd_, args, z, c, alpha, tex_real, mesh_real, training=True):
    with torch.no_grad():
        tex, mesh = generator(sg, args, z, c, training)
        x = torch.cat((torch.cat((tex * alpha, alpha), 1), torch.cat((tex_real, alpha), 1)), 0)
        m = torch.cat((mesh, mesh_real), 0)
    out, mask = discriminator(sd_, args, x, m, torch.cat((c, c), 0), training)
    B = z.shape[0]
    lf = hinge([o[:B] for o in out], False, True, [k[:B] for k in mask])
    lr = hinge([o[B:] for o in out], True, True, [k[B:] for k in mask])
    return lf, lr, out


def trainable(sd):
    """names of the tensors a module would expose as parameters (everything float except buffers)."""
    return [k for k, v in sd.items() if torch.is_floating_point(v) and not k.endswith(("_u", "_v", "running_mean", "running_var"))]


def init_state(args, seed=0):
    """Random initial state dicts (generator, discriminators) with the reference's key names and shapes, built from
    plain tensors — so that the CPU arms of bench.py never import the product package (models.gan loads libb3d.so).
    Shapes follow models/gan.py:23-121,123-233,288-312,314-426 of the reference; values: N(0, 1/fan_in) weights, zero
    biases, unit-norm spectral-norm vectors, fresh batch-norm buffers (enough for a timing baseline and for parity
    tests that load explicit weights afterwards)."""
    g = torch.Generator().manual_seed(seed)
    R, nd = args.texture_resolution, args.num_discriminators

    def w(*shape):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(*shape, generator=g) / math.sqrt(fan_in)

    def sn(sd, name, co, ci, k, bias):
        sd[name + ".weight_orig"] = w(co, ci, k, k)
        if bias:
            sd[name + ".bias"] = torch.zeros(co)
        sd[name + ".weight_u"] = F.normalize(torch.randn(co, generator=g), dim=0)
        sd[name + ".weight_v"] = F.normalize(torch.randn(ci * k * k, generator=g), dim=0)

    def cbn_(sd, name, ch, emb):
        sd[name + ".norm.running_mean"], sd[name + ".norm.running_var"] = torch.zeros(ch), torch.ones(ch)
        sd[name + ".norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        for fc in ("fc_gamma", "fc_beta"):
            sd[f"{name}.{fc}.weight"], sd[f"{name}.{fc}.bias"] = w(ch, emb), torch.zeros(ch)

    def block(sd, name, ci, co, emb):
        mid = min(ci, co)
        sn(sd, name + ".conv1", mid, ci, 3, False)
        sn(sd, name + ".conv2", co, mid, 3, False)
        cbn_(sd, name + ".norm1", mid, emb)
        cbn_(sd, name + ".norm2", co, emb)
        if ci != co:
            sn(sd, name + ".shortcut", co, ci, 1, False)

    sg, emb = {}, 128
    sg["emb_class.weight"] = torch.randn(args.n_classes[0], 64, generator=g)
    sg["fc.weight"], sg["fc.bias"] = w(16384, emb), torch.zeros(16384)
    block(sg, "blk1", 512, 512, emb)
    block(sg, "blk2", 512, 256, emb)
    for name, need in (("blk3a", 256), ("blk3b", 512), ("blk3c", 1024)):
        if R >= need:
            block(sg, name, 256, 256, emb)
    block(sg, "blk4", 256, 128, emb)
    block(sg, "blk5", 128, 128, emb)
    block(sg, "blk6", 128, 64, emb)
    sg["conv_final.weight"], sg["conv_final.bias"] = w(3, 64, 5, 5), torch.zeros(3)
    block(sg, "blk3_mesh", 256, 64, emb)
    sg["conv_mesh.weight"], sg["conv_mesh.bias"] = w(3, 64, 5, 5) * 0.02, torch.zeros(3)

    sd = {}

    def texd(pre, downsample):
        stride_first = (downsample == 1 and R >= 512) or R >= 1024
        sn(sd, pre + "conv1", 64, 8, 4 if stride_first else 5, True)
        sn(sd, pre + "conv2", 128, 64, 4, True)
        sn(sd, pre + "conv3", 256, 128, 4, True)
        sn(sd, pre + "conv4", 512, 256, 4, True)
        sn(sd, pre + "conv5", 1, 512, 5, True)
        sd[pre + "projector.weight"] = torch.randn(args.n_classes[0], 512, generator=g)

    texd("d1.", 1)
    sn(sd, "d2.conv1", 64, 11, 5, True)
    sn(sd, "d2.conv2", 128, 64, 4, True)
    sn(sd, "d2.conv3", 256, 128, 4, True)
    sn(sd, "d2.conv4", 1, 256, 5, True)
    sd["d2.projector.weight"] = torch.randn(args.n_classes[0], 256, generator=g)
    if nd == 3:
        texd("d3.", 4)
    return sg, sd
