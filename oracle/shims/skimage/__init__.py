"""Import stand-in for scikit-image (not installed here); only metrics.peak_signal_noise_ratio is used
by the reference (encoder.py:9,123; Frame.py:6; decoder.py:5).  TEST INFRASTRUCTURE ONLY."""
