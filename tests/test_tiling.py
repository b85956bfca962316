"""Tiled full-frame inference (BASELINE config 4).  CPU: tile grid / weights / blend restatement properties.
GPU: gather + blend kernels against the restatement, whole tiled pass against per-tile CPU oracle forwards + the
restated blend, and the 1080x1920 / 256-tile configuration through size-independent properties."""
import pytest
import torch

from image_restoration_b200.tiling import blend_reference, ramp_weights, tile_positions


def test_tile_grid_matches_config4():
    assert tile_positions(1080, 256, 32) == [0, 224, 448, 672, 824]
    assert len(tile_positions(1920, 256, 32)) == 9 and tile_positions(1920, 256, 32)[-1] == 1920 - 256
    assert tile_positions(256, 256, 32) == [0]
    with pytest.raises(ValueError):
        tile_positions(200, 256, 32)
    w = ramp_weights(1080, 256, 32, tile_positions(1080, 256, 32))
    assert w.shape == (5, 256) and (w > 0).all() and w[0, 0] == 1 and w[-1, -1] == 1
    assert abs(w[1, 0].item() - 1 / 33) < 1e-7 and w[1, 31].item() < 1 and w[1, 32].item() == 1


def test_blend_is_a_partition_of_unity_and_identity_on_consistent_tiles():
    torch.manual_seed(0)
    H, W, T, ov = 150, 200, 64, 16
    frame = torch.rand(3, H, W)
    ty, tx = tile_positions(H, T, ov), tile_positions(W, T, ov)
    tiles = torch.stack([frame[:, y:y + T, x:x + T] for y in ty for x in tx])
    assert torch.allclose(blend_reference(tiles, H, W, T, ov), frame, atol=1e-6)       # consistent tiles -> the frame
    assert torch.allclose(blend_reference(torch.full_like(tiles, 0.7), H, W, T, ov), torch.full((3, H, W), 0.7), atol=1e-6)


@pytest.mark.gpu
def test_gather_and_blend_kernels_match_restatement():
    from image_restoration_b200 import _lib
    from image_restoration_b200.ops import _ptr, _stream
    torch.manual_seed(1)
    for (H, W, T, ov) in [(150, 200, 64, 16), (1080, 1920, 256, 32), (64, 64, 64, 8)]:
        frame = torch.rand(3, H, W, device='cuda')
        ty, tx = tile_positions(H, T, ov), tile_positions(W, T, ov)
        ty_d = torch.tensor(ty, dtype=torch.int32, device='cuda')
        tx_d = torch.tensor(tx, dtype=torch.int32, device='cuda')
        n = len(ty) * len(tx)
        tiles = torch.empty(n, 3, T, T, device='cuda')
        lib = _lib.lib()
        _lib.check(lib.b200ir_tiles_gather(_ptr(frame), _ptr(tiles), 3, H, W, T, _ptr(ty_d), _ptr(tx_d), len(ty), len(tx),
                                           _stream()))
        ref_tiles = torch.stack([frame[:, y:y + T, x:x + T] for y in ty for x in tx])
        assert torch.equal(tiles, ref_tiles)
        noisy = tiles + 0.1 * torch.randn_like(tiles)                                    # inconsistent tiles
        out = torch.empty_like(frame)
        _lib.check(lib.b200ir_tiles_blend(_ptr(noisy), _ptr(out), 3, H, W, T, ov, _ptr(ty_d), _ptr(tx_d), len(ty), len(tx),
                                          _stream()))
        ref = blend_reference(noisy.cpu(), H, W, T, ov)
        assert (out.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_tiled_pass_matches_per_tile_oracle():
    """Small geometry the CPU oracle finishes in seconds: 64x64 network, 150x200 frame, overlap 16 (3x4 = 12 tiles)."""
    from image_restoration_b200 import GFPGANv1OCR
    from image_restoration_b200.tiling import TiledRestorer
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward, psnr01, to01
    torch.manual_seed(2)
    kw = dict(input_width=64, input_height=64, num_style_feat=256, channel_multiplier=0.5, num_mlp=4,
              input_is_latent=True, different_w=True, narrow=1, sft_half=True)
    net = GFPGANv1OCR(decoder_load_path=None, fix_decoder=True, **kw).eval()
    frame = torch.rand(3, 150, 200) * 2 - 1
    ty, tx = tile_positions(150, 64, 16), tile_positions(200, 64, 16)
    tiles = torch.stack([frame[:, y:y + 64, x:x + 64] for y in ty for x in tx])
    ref_tiles, _ = gfpgan_ocr_forward(net.state_dict(), OcrNetConfig(**kw), tiles, False)
    ref = blend_reference(ref_tiles, 150, 200, 64, 16)
    got = TiledRestorer(net.cuda(), overlap=16)(frame.cuda(), randomize_noise=False).cpu()
    a, b = to01(got), to01(ref)
    max_abs, psnr = (a - b).abs().max().item(), psnr01(a, b)
    print(f'tiled 150x200 / 64: max-abs {max_abs:.3e} psnr {psnr:.1f} dB')
    assert max_abs <= 2e-2 and psnr >= 45.0


@pytest.mark.gpu
def test_config4_full_frame_properties():
    """3x1080x1920 frame, 256x256 tiles, overlap 32 -> 45 tiles in one batch: output shape / finiteness, and the result
    equals blending the network's own per-tile outputs (tile order and weights are the only things the tiler adds)."""
    from image_restoration_b200 import GFPGANv1OCR
    from image_restoration_b200.tiling import TiledRestorer
    torch.manual_seed(3)
    kw = dict(input_width=256, input_height=256, num_style_feat=256, channel_multiplier=0.5, num_mlp=4,
              input_is_latent=True, different_w=True, narrow=1, sft_half=True)
    net = GFPGANv1OCR(decoder_load_path=None, fix_decoder=True, **kw).eval().cuda()
    frame = torch.rand(3, 1080, 1920, device='cuda')
    tiler = TiledRestorer(net, overlap=32, micro_batch=45)
    out = tiler(frame, randomize_noise=False)
    assert out.shape == frame.shape and torch.isfinite(out).all()
    ty, tx = tile_positions(1080, 256, 32), tile_positions(1920, 256, 32)
    assert len(ty) * len(tx) == 45
    tiles = torch.stack([frame[:, y:y + 256, x:x + 256] for y in ty for x in tx])
    per_tile = net(tiles, return_rgb=False, randomize_noise=False)[0].cpu()
    ref = blend_reference(per_tile, 1080, 1920, 256, 32)
    assert (out.cpu() - ref).abs().max().item() < 1e-4
