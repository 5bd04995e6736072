"""weatherbench2_b200 -- B200-native hot path of WeatherBench 2.

Per-chunk evaluation metrics, conservative regridding and the zonal energy
spectrum, behind the reference's own Python operator API
(`Metric.compute_chunk`, `Region.apply`, `evaluation._metric_and_region_loop`,
`ConservativeRegridder.regrid_array`, `ZonalEnergySpectrum.compute`), with the
arithmetic in hand-written CUDA for sm_100a reached through a C ABI
(`include/wb2b200.h`, `libwb2b200.so`).  No CPU fallback.
"""
__version__ = '0.1.0'

# Import order matters for the private implementation modules (`_ensemble`,
# `_thresholded`, `_seeps`, `_rank_hist`, `_regrid_interp`, `_derived_wind`): they
# extend the public modules below, which import them at their end.  Loading the
# public modules here makes `from weatherbench2_b200 import _seeps` (or any other
# private module) work in any order.  Nothing here touches the GPU or loads
# libwb2b200.so.
from weatherbench2_b200 import metrics  # noqa: E402,F401  pylint: disable=wrong-import-position
from weatherbench2_b200 import regridding  # noqa: E402,F401  pylint: disable=wrong-import-position
from weatherbench2_b200 import derived_variables  # noqa: E402,F401  pylint: disable=wrong-import-position
