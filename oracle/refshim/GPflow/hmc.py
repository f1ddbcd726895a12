"""GPflow.hmc.sample_HMC (restated): leapfrog HMC over the free state; f returns (-logp, -grad)."""
import numpy as np


def sample_HMC(f, num_samples, Lmax, epsilon, x0, verbose=False, thin=1, burn=0, RNG=np.random.RandomState(0),
               return_logprobs=False):
    if burn > 0:
        samples = sample_HMC(f, num_samples=burn, Lmax=Lmax, epsilon=epsilon, x0=x0, verbose=verbose, thin=1, burn=0,
                             RNG=RNG)
        x0 = samples[-1]
    D = x0.size
    samples = np.zeros((num_samples, D))
    logprobs = np.zeros(num_samples)
    x = x0.copy()
    logprob, grad = f(x0)
    logprob, grad = -logprob, -grad
    for t in range(num_samples):
        for _ in range(thin):
            x_old, logprob_old, grad_old = x.copy(), logprob, grad.copy()
            p_old = RNG.randn(D)
            premature_reject = False
            p = p_old + 0.5 * epsilon * grad
            for l in range(RNG.randint(1, Lmax)):
                x = x + epsilon * p
                logprob, grad = f(x)
                logprob, grad = -logprob, -grad
                if np.any(np.isnan(grad)):
                    premature_reject = True
                    break
                p = p + epsilon * grad
            p = p - 0.5 * epsilon * grad
            if premature_reject:
                x, logprob, grad = x_old, logprob_old, grad_old
                continue
            log_accept_ratio = logprob - 0.5 * p.dot(p) - logprob_old + 0.5 * p_old.dot(p_old)
            if np.log(RNG.rand()) >= log_accept_ratio:
                x, logprob, grad = x_old, logprob_old, grad_old
        samples[t] = x
        logprobs[t] = logprob
    return (samples, logprobs) if return_logprobs else samples
