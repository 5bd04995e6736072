"""weatherbench2_b200 -- B200-native hot path of WeatherBench 2.

Per-chunk evaluation metrics, conservative regridding and the zonal energy
spectrum, behind the reference's own Python operator API
(`Metric.compute_chunk`, `Region.apply`, `evaluation._metric_and_region_loop`,
`ConservativeRegridder.regrid_array`, `ZonalEnergySpectrum.compute`), with the
arithmetic in hand-written CUDA for sm_100a reached through a C ABI
(`include/wb2b200.h`, `libwb2b200.so`).  No CPU fallback.
"""
__version__ = '0.1.0'
