"""Import-time stand-in for xarray_beam (see ../apache_beam)."""


class Key:
  pass
