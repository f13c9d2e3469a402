"""PlotBER curve store (my_sn/plotting.py:18-48).  matplotlib is optional (absent in the build image):
simulate() never needs it, plot_ber() imports it lazily."""
from .sim import sim_ber


def plot_ber(plot_self, ylabel="BER"):
  import matplotlib.pyplot as plt
  fig, ax = plt.subplots(figsize=(16, 10))
  plt.xticks(fontsize=18); plt.yticks(fontsize=18)
  plt.title(plot_self.title, fontsize=25)
  for idx, b in enumerate(plot_self.ber):
    plt.semilogy(plot_self.snr[idx], b, linewidth=2)
  plt.grid(which="both"); plt.xlabel(r"$E_b/N_0$ (dB)", fontsize=25)
  plt.ylabel(ylabel, fontsize=25); plt.legend(plot_self.legend, fontsize=20)
  return fig, ax


class PlotBER():
  def __init__(self, title="Bit/Block Error Rate"):
    self.title = title
    self.ber = []; self.snr = []; self.legend = []

  def simulate(self, mc_fun, ebno_dbs, batch_size, legend="", add_ber=True, add_bler=False, max_mc_iter=1,
               soft_estimates=False, target_bit_errs=None, target_block_errs=None, verbose=True, device='cpu',
               on_device=True):
    ber, bler = sim_ber(mc_fun, ebno_dbs, batch_size, soft_estimates=soft_estimates, max_mc_iter=max_mc_iter,
                        target_bit_errs=target_bit_errs, target_block_errs=target_block_errs, verbose=verbose,
                        device=device, on_device=on_device)
    if add_ber:
      self.ber += [ber]; self.snr += [ebno_dbs]; self.legend += [legend]
    if add_bler:
      self.ber += [bler]; self.snr += [ebno_dbs]; self.legend += [legend + " (BLER)"]
    return ber, bler
