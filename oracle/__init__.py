"""CPU oracle for the VDO-SLAM hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  Nothing under vdo_slam_b200/ does.
"""
