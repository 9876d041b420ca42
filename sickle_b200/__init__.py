"""sickle_b200 -- B200-native trimming hot path of Parallel Sickle."""
