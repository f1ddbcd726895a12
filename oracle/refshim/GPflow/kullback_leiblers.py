"""GPflow.kullback_leiblers.gauss_kl_white[_diag]: KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, I)] summed
over the latent columns."""
import tensorflow as tf
from ._settings import settings
float_type = settings.dtypes.float_type


def gauss_kl_white(q_mu, q_sqrt):
    KL = 0.5 * tf.reduce_sum(tf.square(q_mu))                     # Mahalanobis term
    KL += -0.5 * tf.cast(tf.size(q_mu), float_type)               # constant term
    L = tf.batch_matrix_band_part(tf.transpose(q_sqrt, (2, 0, 1)), -1, 0)
    KL -= 0.5 * tf.reduce_sum(tf.log(tf.square(tf.batch_matrix_diag_part(L))))   # log-det of q-cov
    KL += 0.5 * tf.reduce_sum(tf.square(L))                       # trace term
    return KL


def gauss_kl_white_diag(q_mu, q_sqrt):
    KL = 0.5 * tf.reduce_sum(tf.square(q_mu))
    KL += -0.5 * tf.cast(tf.size(q_sqrt), float_type)
    KL += -0.5 * tf.reduce_sum(tf.log(tf.square(q_sqrt)))
    KL += 0.5 * tf.reduce_sum(tf.square(q_sqrt))
    return KL
