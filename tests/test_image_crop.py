"""image_crop: image_crop::ImageCrop::cameraCallback (image_crop/src/image_crop.cpp:20-43).

CPU: the oracle restatement against the REFERENCE'S OWN image_crop.cpp (compiled unmodified into
oracle/_ref/libmod_ref.so against the stand-in dependency headers) and the C ABI's host-side CameraInfo fix-up against
both.  GPU (-m gpu): the crop kernel through the C ABI (host seam and device-resident batch) against the oracle and the
reference build."""
import numpy as np
import pytest

CHANNELS = {"bgr8": 3, "rgb8": 3, "bgra8": 4, "rgba8": 4, "mono8": 1}
#            w   h   tw  th
SIZES = [(64, 48, 32, 20), (65, 47, 33, 21), (40, 30, 40, 30), (17, 9, 1, 1), (17, 9, 0, 0), (33, 31, 32, 31),
         (40, 30, 41, 30), (40, 30, 10, 31), (5, 5, 200, 2)]   # the last three windows do not fit


def _image(rng, w, h, enc, pad=0):
    ch = CHANNELS[enc]
    full = rng.randint(0, 256, (h, w + pad, ch) if ch > 1 else (h, w + pad)).astype(np.uint8)
    return full[:, :w]   # pad > 0: a strided view, image.step > width * channels


def _caminfo(w, h):
    K = np.array([500.25, 0, w / 2 + 0.3, 0, 501.5, h / 2 - 0.2, 0, 0, 1.0])
    P = np.array([500.25, 0, w / 2 + 0.3, -35.0, 0, 501.5, h / 2 - 0.2, 0.5, 0, 0, 1.0, 0.0])
    return K, P


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not built (needs /root/reference at build time)")
    ref_py.load()
    return ref_py


@pytest.mark.parametrize("enc", sorted(CHANNELS))
def test_oracle_equals_reference_image_crop(oracle, ref, enc):
    rng = np.random.RandomState(11)
    for w, h, tw, th in SIZES:
        for pad in (0, 3):
            img = _image(rng, w, h, enc, pad)
            K, P = _caminfo(w, h)
            r = ref.image_crop(img, enc, tw, th, K, P)
            o = oracle.image_crop(img, enc, tw, th)
            if o is None:
                assert r["status"] == -12, "the reference's ROI assertion must throw where the oracle rejects"
                continue
            assert r["status"] == 0
            assert np.array_equal(o, r["image"]), (enc, w, h, tw, th, pad)
            Ko, Po = oracle.crop_camera_info(w, h, tw, th, K, P)
            assert np.array_equal(Ko, r["K"]) and np.array_equal(Po, r["P"])
            assert (r["width"], r["height"]) == (tw, th) and r["image_wh_step"] == (tw, th, tw * 3)


def test_host_camera_info_fixup_equals_reference(oracle, ref, mod):
    from moving_object_detector_b200 import ModError
    from moving_object_detector_b200.api import crop_camera_info
    rng = np.random.RandomState(12)
    for w, h, tw, th in SIZES:
        K, P = _caminfo(w, h)
        img = _image(rng, w, h, "mono8")
        r = ref.image_crop(img, "mono8", tw, th, K, P)
        if r["status"] != 0:
            with pytest.raises(ModError):
                crop_camera_info(w, h, tw, th, K, P)
            continue
        Ko, Po = crop_camera_info(w, h, tw, th, K, P)
        assert np.array_equal(Ko, r["K"]) and np.array_equal(Po, r["P"])
        # odd differences round down: (65 - 33) / 2 = 16, (47 - 21) / 2 = 13 (unsigned integer division)
        assert Ko[2] == K[2] - (w - tw) // 2 and Ko[5] == K[5] - (h - th) // 2 and Po[3] == P[3] and Po[7] == P[7]


# ---- GPU -----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from moving_object_detector_b200 import Context
    c = Context(1280, 720, max_batch=1)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("enc", sorted(CHANNELS))
def test_gpu_crop_equals_oracle_and_reference(ctx, oracle, enc):
    import ref_py
    from moving_object_detector_b200 import ModError
    rng = np.random.RandomState(21)
    for w, h, tw, th in SIZES + [(1280, 720, 1024, 576), (1279, 719, 1001, 333)]:
        for pad in (0, 5):
            img = _image(rng, w, h, enc, pad)
            want = oracle.image_crop(img, enc, tw, th)
            if want is None:
                with pytest.raises(ModError):
                    ctx.image_crop(img, enc, tw, th)
                continue
            got = ctx.image_crop(img, enc, tw, th)
            assert np.array_equal(got, want), (enc, w, h, tw, th, pad)
            if ref_py.available():
                K, P = _caminfo(w, h)
                assert np.array_equal(got, ref_py.image_crop(img, enc, tw, th, K, P)["image"])


@pytest.mark.gpu
def test_gpu_crop_device_batch(oracle):
    """Device-resident batch of ZED-like frames (detect_with_zed.launch crops both cameras before SGM / PWC-Net)."""
    import torch
    from moving_object_detector_b200 import Context
    rng = np.random.RandomState(22)
    w, h, tw, th, batch = 1280, 720, 1024, 512, 6
    frames = rng.randint(0, 256, (batch, h, w, 4)).astype(np.uint8)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((batch, th, tw, 3), dtype=torch.uint8, device="cuda")
    with Context(w, h, max_batch=1) as c:
        c.image_crop_batched_dev(w, h, batch, "bgra8", d_in, h * w * 4, w * 4, tw, th, d_out,
                                 stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    for i in range(batch):
        assert np.array_equal(got[i], oracle.image_crop(frames[i], "bgra8", tw, th)), i
