"""B200-native drop-in for the reference ``sesameai`` package (CSM-1B frame generation hot path)."""
