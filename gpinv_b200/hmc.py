"""Hamiltonian Monte Carlo over the packed free state (GPflow ``hmc.sample_HMC``, the sampler
behind ``GPMC.sample`` at notebooks/Abel_inversion.ipynb:593).  ``f(x)`` returns
(-log density, -gradient) — the ``_objective`` convention.  Per iteration: momentum draw,
``RNG.randint(1, Lmax)`` leapfrog steps, NaN gradient => premature reject, Metropolis test."""
import numpy as np


def sample_HMC(f, num_samples, Lmax, epsilon, x0, verbose=False, thin=1, burn=0,
               RNG=None, return_logprobs=False):
    if RNG is None:
        RNG = np.random.RandomState(0)
    if burn > 0:
        samples = sample_HMC(f, num_samples=burn, Lmax=Lmax, epsilon=epsilon, x0=x0,
                             verbose=verbose, thin=1, burn=0, RNG=RNG)
        if verbose:
            print("burn-in sampling ended")
        x0 = samples[-1]

    D = x0.size
    samples = np.zeros((num_samples, D))
    logprobs = np.zeros(num_samples)
    x = x0.copy()
    logprob, grad = f(x0)
    logprob, grad = -logprob, -grad
    accept_count_batch = 0
    for t in range(num_samples):
        for t_ in range(thin):
            it = t * thin + t_
            if it % 100 == 0 and it > 0:
                if verbose:
                    print("Iteration: ", it, "\t Acc Rate: ", 1. * accept_count_batch, "%")
                accept_count_batch = 0
            x_old, logprob_old, grad_old = x.copy(), logprob, grad.copy()
            p_old = RNG.randn(D)
            premature_reject = False
            p = p_old + 0.5 * epsilon * grad
            for _ in range(RNG.randint(1, Lmax)):
                x = x + epsilon * p
                logprob, grad = f(x)
                logprob, grad = -logprob, -grad
                if np.any(np.isnan(grad)) or np.isnan(logprob):
                    premature_reject = True
                    break
                p = p + epsilon * grad
            p = p - 0.5 * epsilon * grad
            if premature_reject:
                if verbose:
                    print("warning: numerical instability. Rejecting this proposal prematurely")
                x, logprob, grad = x_old, logprob_old, grad_old
                continue
            log_accept_ratio = logprob - 0.5 * p.dot(p) - logprob_old + 0.5 * p_old.dot(p_old)
            logu = np.log(RNG.rand())
            if logu < log_accept_ratio:
                accept_count_batch += 1
            else:
                x, logprob, grad = x_old, logprob_old, grad_old
        samples[t] = x
        logprobs[t] = logprob
    if return_logprobs:
        return samples, logprobs
    return samples


def sample_HMC_device(f, num_samples, Lmax, epsilon, x0, device, verbose=False, thin=1, burn=0,
                      RNG=None, return_logprobs=False):
    """The same chain as ``sample_HMC`` (same RNG consumption, same accept rule) with the state,
    momentum, gradient and the sample matrix resident on the device.  ``f(x)`` takes / returns
    CUDA tensors (``Model._objective_device``: ([1], [D]) = (-log density, -gradient); it may be a
    replayed CUDA graph, so its outputs are consumed before the next call).  The leapfrog steps
    never touch the host; there is ONE small device->host read per HMC iteration (the NaN flag
    and the log acceptance ratio), which the reference's control flow needs because a premature
    reject skips the uniform draw.  SURVEY.md 8(f) rank 2 (loop shape: GPflow hmc.sample_HMC)."""
    import torch
    if RNG is None:
        RNG = np.random.RandomState(0)
    x0 = np.asarray(x0, dtype=np.float64)
    if burn > 0:
        s = sample_HMC_device(f, burn, Lmax, epsilon, x0, device, verbose=verbose, thin=1, burn=0, RNG=RNG)
        if verbose:
            print("burn-in sampling ended")
        x0 = s[-1]
    D = x0.size
    samples = torch.zeros((num_samples, D), dtype=torch.float64, device=device)
    logprobs = torch.zeros(num_samples, dtype=torch.float64, device=device)
    x = torch.as_tensor(x0).to(device)

    def evaluate(xx):
        nf, ng = f(xx)
        return -nf.reshape(()), -ng          # new tensors: the graph's static outputs are free again

    logprob, grad = evaluate(x)
    accept_count_batch = 0
    for t in range(num_samples):
        for t_ in range(thin):
            it = t * thin + t_
            if it % 100 == 0 and it > 0:
                if verbose:
                    print("Iteration: ", it, "\t Acc Rate: ", 1. * accept_count_batch, "%")
                accept_count_batch = 0
            x_old, logprob_old, grad_old = x, logprob, grad
            p_old = torch.as_tensor(RNG.randn(D)).to(device)
            p = p_old + 0.5 * epsilon * grad
            bad = torch.zeros((), dtype=torch.bool, device=device)
            for _ in range(RNG.randint(1, Lmax)):
                x = x + epsilon * p
                logprob, grad = evaluate(x)
                bad = bad | torch.isnan(logprob) | torch.isnan(grad).any()
                p = p + epsilon * grad
            p = p - 0.5 * epsilon * grad
            log_accept_ratio = logprob - 0.5 * p.dot(p) - logprob_old + 0.5 * p_old.dot(p_old)
            flags = torch.stack([bad.to(torch.float64), log_accept_ratio]).cpu()     # the one host read
            if flags[0].item() != 0.0:
                if verbose:
                    print("warning: numerical instability. Rejecting this proposal prematurely")
                x, logprob, grad = x_old, logprob_old, grad_old
                continue
            logu = np.log(RNG.rand())
            if logu < flags[1].item():
                accept_count_batch += 1
            else:
                x, logprob, grad = x_old, logprob_old, grad_old
        samples[t] = x
        logprobs[t] = logprob
    if return_logprobs:
        return samples.cpu().numpy(), logprobs.cpu().numpy()
    return samples.cpu().numpy()
