"""Architecture of the hot path as shipped by the reference (`config/models/visinger.yaml:22-29`,
`models/visinger.py:65-69`): constructor arguments of the flow and the HiFi-GAN generator, in the short-key form the
module mirrors' `from_configs` / `random_init` helpers take."""

VISINGER_FLOW = dict(channels=192, hidden=192, kernel_size=5, dilation_rate=1, n_layers=4, n_flows=4, gin=256)
VISINGER_GENERATOR = dict(initial_channel=192, resblock="1", rk=[3, 7, 11], rd=[[1, 3, 5]] * 3, ur=[5, 5, 3, 2, 2], uic=512,
                          uk=[11, 11, 7, 4, 4], gin=256)
