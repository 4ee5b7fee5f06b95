"""btrc_b200: B200-native drop-in for the wavefront render path of AirGuanZ/Btrc."""
