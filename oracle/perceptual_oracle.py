"""ORACLE (test infrastructure, NOT product code): fp32 torch restatement of the reference's perceptual loss.

  PerceptualLoss.forward / _gram_mat   Car_Plate-Restoration/basicsr/losses/losses.py:250-356
  VGGFeatureExtractor.forward          Car_Plate-Restoration/basicsr/archs/vgg_arch.py:56-160  (range_norm, use_input_norm,
                                       torchvision vgg19 `features` up to the last tapped layer, taps read BEFORE their ReLU)

The VGG itself is third-party (torchvision.models.vgg19, pinned torchvision==0.14.0 in requirements.txt; 0.26 here): a stack
of nn.Conv2d(3x3, padding 1) + ReLU with nn.MaxPool2d(2, 2) after conv1_2 / 2_2 / 3_4 / 4_4 / 5_4.  It is restated functionally
over a state_dict with torchvision's key names (`features.<idx>.weight|bias`) so that it runs without torchvision; pinned
against torchvision's module on the same weights in tests/test_perceptual_cpu.py.  The reference loads ImageNet weights
(vgg19-dcbb9e9d.pth / VGG19_Weights.DEFAULT) — not available offline: tests use seeded random weights; weights are data.
"""
import torch
import torch.nn.functional as F

# torchvision vgg19.features: index of every conv and the reference's layer names (vgg_arch.py:34-40)
VGG19_NAMES = ['conv1_1', 'relu1_1', 'conv1_2', 'relu1_2', 'pool1', 'conv2_1', 'relu2_1', 'conv2_2', 'relu2_2', 'pool2',
               'conv3_1', 'relu3_1', 'conv3_2', 'relu3_2', 'conv3_3', 'relu3_3', 'conv3_4', 'relu3_4', 'pool3', 'conv4_1',
               'relu4_1', 'conv4_2', 'relu4_2', 'conv4_3', 'relu4_3', 'conv4_4', 'relu4_4', 'pool4', 'conv5_1', 'relu5_1',
               'conv5_2', 'relu5_2', 'conv5_3', 'relu5_3', 'conv5_4', 'relu5_4', 'pool5']
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def vgg_features(sd, x, layer_names, use_input_norm=True, range_norm=True):
    """vgg_arch.py:140-160."""
    if range_norm:
        x = (x + 1) / 2
    if use_input_norm:
        x = (x - torch.tensor(MEAN).view(1, 3, 1, 1).to(x)) / torch.tensor(STD).view(1, 3, 1, 1).to(x)
    out = {}
    last = max(VGG19_NAMES.index(n) for n in layer_names)
    for idx, name in enumerate(VGG19_NAMES[:last + 1]):
        if name.startswith('conv'):
            x = F.conv2d(x, sd[f'features.{idx}.weight'].to(x), sd[f'features.{idx}.bias'].to(x), padding=1)
        elif name.startswith('relu'):
            x = F.relu(x)
        else:
            x = F.max_pool2d(x, 2, 2)
        if name in layer_names:
            out[name] = x.clone()
    return out


def gram_mat(x):
    """losses.py:343-356."""
    n, c, h, w = x.shape
    f = x.view(n, c, w * h)
    return f.bmm(f.transpose(1, 2)) / (c * h * w)


def perceptual_loss(sd, x, gt, layer_weights, perceptual_weight=1.0, style_weight=0.0, use_input_norm=True, range_norm=True):
    """losses.py:300-341 with criterion='l1'.  Returns (percep_loss | None, style_loss | None)."""
    fx = vgg_features(sd, x, list(layer_weights), use_input_norm, range_norm)
    fg = vgg_features(sd, gt.detach(), list(layer_weights), use_input_norm, range_norm)
    percep = style = None
    if perceptual_weight > 0:
        percep = sum(F.l1_loss(fx[k], fg[k]) * layer_weights[k] for k in fx) * perceptual_weight
    if style_weight > 0:
        style = sum(F.l1_loss(gram_mat(fx[k]), gram_mat(fg[k])) * layer_weights[k] for k in fx) * style_weight
    return percep, style


def random_vgg19_state_dict(seed=0):
    """Seeded stand-in for the ImageNet checkpoint: Kaiming-normal conv weights (torchvision's own init), small random biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 3
    for idx, name in enumerate(VGG19_NAMES):
        if not name.startswith('conv'):
            continue
        cout = {'1': 64, '2': 128, '3': 256, '4': 512, '5': 512}[name[4]]
        sd[f'features.{idx}.weight'] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cout * 9)) ** 0.5
        sd[f'features.{idx}.bias'] = torch.randn(cout, generator=g) * 0.05
        cin = cout
    return sd
