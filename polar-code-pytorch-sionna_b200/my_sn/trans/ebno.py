def ebnodb2no(ebno_db, n_bits_per_sym, coderate):
  """No = 1 / (Eb/No * coderate * bits_per_symbol), Es = 1 (my_sn/trans/ebno.py:21-23)."""
  ebno = 10. ** (ebno_db / 10.)
  energy_per_symbol = 1
  return 1 / (ebno * coderate * n_bits_per_sym / energy_per_symbol)
