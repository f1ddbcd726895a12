"""GPflow._settings: float_type = float64, numerics.jitter_level = 1e-6 (gpflowrc defaults)."""
import tensorflow as tf


class _NS(object):
    pass


settings = _NS()
settings.dtypes = _NS()
settings.dtypes.float_type = tf.float64
settings.dtypes.int_type = tf.int32
settings.numerics = _NS()
settings.numerics.jitter_level = 1e-6
