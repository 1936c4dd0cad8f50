"""shared test helpers (importable because pytest puts tests/ on sys.path)"""
import numpy as np

from deepfactors_b200 import synth


def scenenet_inputs(golden):
    """inputs of tests/ut_se3aligner.cpp:58-77: 1047 -> 1052, /255, 25x25 box blur, depth mm -> m, SceneNet camera"""
    img0 = golden["blur25_1047"]
    img1 = golden["blur25_1052"]
    dpt0 = (golden["depth_1047_mm"].astype(np.float32) * np.float32(1 / 1000.0)).astype(np.float32)
    cam = synth.Camera.scenenet(img0.shape[1], img0.shape[0])
    return cam, img0, img1, dpt0
