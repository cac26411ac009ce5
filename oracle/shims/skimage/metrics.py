import numpy as np


def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    """10*log10(255^2 / mse) for uint8 inputs (skimage definition)."""
    a = np.asarray(image_true, dtype=np.float64)
    b = np.asarray(image_test, dtype=np.float64)
    if data_range is None:
        data_range = 255
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    return 10 * np.log10((data_range ** 2) / mse)
