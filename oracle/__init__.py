"""CPU oracle package — TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product (mirror_maze_b200/) never does."""
