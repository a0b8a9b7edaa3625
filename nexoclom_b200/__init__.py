"""nexoclom_b200 -- B200-native hot path of the nexoclom exosphere model."""
