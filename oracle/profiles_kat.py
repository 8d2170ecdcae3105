"""Chemistry profile constants restated from utils.py:149-210 (TEST INFRASTRUCTURE: the oracle's own
copy, checked against tests/golden/profiles.json which was dumped from the reference's get_profile)."""
PROFILES = {
    "dna-r10-min": dict(digitisation=8192, sample_rate=5000, bps=400, range=1536.598389,
                        offset_mean=13.380569389019, offset_std=16.311471649012,
                        median_before_mean=202.15407438804, median_before_std=13.406139241768),
    "dna-r10-prom": dict(digitisation=2048, sample_rate=5000, bps=400, range=281.345551,
                         offset_mean=-127.5655735, offset_std=19.377283387665,
                         median_before_mean=189.87607393756, median_before_std=15.788097978713),
    "dna-r9-min": dict(digitisation=8192, sample_rate=4000, bps=450, range=1443.030273,
                       offset_mean=13.7222605, offset_std=10.25279688,
                       median_before_mean=200.815801, median_before_std=20.48933762),
    "dna-r9-prom": dict(digitisation=2048, sample_rate=4000, bps=450, range=748.5801,
                        offset_mean=-237.4102, offset_std=14.1575,
                        median_before_mean=214.2890337, median_before_std=18.0127916),
    "rna-004-min": dict(digitisation=8192, sample_rate=4000, bps=130, range=1437.976685,
                        offset_mean=12.47686423863, offset_std=10.442126577137,
                        median_before_mean=205.08496731088, median_before_std=8.6671292866233),
    "rna-004-prom": dict(digitisation=2048, sample_rate=4000, bps=130, range=299.432068,
                         offset_mean=-259.421128, offset_std=16.010841823643,
                         median_before_mean=189.87607393756, median_before_std=15.788097978713),
}
